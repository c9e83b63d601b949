for lib in "" meshclust2_b200/lib/variants/lib_k1_4.so meshclust2_b200/lib/variants/lib_k1_6.so; do
  for w in 1 3; do
    echo "== lib=${lib:-default(5)} waves=$w"
    MC2_LIB=$lib MC2_K1_WAVES=$w K1_SHORT=1 K1_NO_LEGACY=1 timeout 60 python tools/k1_bench.py 2>&1 | grep "cfg3 short reads  "
  done
done
