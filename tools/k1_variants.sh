#!/bin/bash
# K1 short-read kernel: register budget (MC2_K1_MIN_CTAS = CTAs of 8 warps per SM) x waves per resident CTA.  Run on a GPU box
# from the repo root.  The variant libraries are not kept; build them first, e.g.
#   for mb in 4 6; do (cd meshclust2_b200 && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 \
#       -Xcompiler -fPIC,-fopenmp,-O2 -shared -lgomp -DMC2_K1_MIN_CTAS=$mb -o lib/variants/lib_k1_$mb.so csrc/*.cu csrc/host_encode.cpp); done
# Round 1: 0.090-0.097 ms per 100k x 1 kb reads for every combination (no sensitivity).
for lib in "" meshclust2_b200/lib/variants/lib_k1_4.so meshclust2_b200/lib/variants/lib_k1_6.so; do
  for w in 1 3; do
    echo "== lib=${lib:-default(5)} waves=$w"
    MC2_LIB=$lib MC2_K1_WAVES=$w K1_SHORT=1 K1_NO_LEGACY=1 timeout 60 python tools/k1_bench.py 2>&1 | grep "cfg3 short reads  "
  done
done
