"""Small fixed workload for ncu captures: one sweep launch (20k x 20k upper triangle, k=5 u8) and two one-vs-many
launches over 2^20 candidates (1 GiB streamed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from meshclust2_b200 import capi

ctx = capi.Context(0)
rng = np.random.default_rng(0)
n = int(os.environ.get("PT_N", 1 << 20))
base = rng.integers(1, 7, size=(4096, 1024), dtype=np.uint8)
H = base[rng.integers(0, 4096, n)]
ln = rng.integers(950, 1050, n).astype(np.uint64)
hs = ctx.hset_from_host(H, 5, length=ln)
gm = ctx.model_from_file(os.path.join("tests", "golden", os.environ.get("PT_W", "weights_cfg1_id90") + ".txt"))
for it in range(int(os.environ.get("PT_ITERS", 3))):
    ms, nc = ctx.bench_score_pairs(gm, hs, hs, n_pairs=n, a_begin=0, b_begin=5, b_bc=1, iters=1, flush_l2=False,
                                   len_filter=1, anchor_is_b=1, cutoff=0.9)
    print("one-vs-many %d: %.3f ms %.3e pairs/s %.1f GB/s close=%d" % (n, ms, n / ms * 1e3, n * 1057 / ms / 1e6, nc))
m = int(os.environ.get("PT_M", 20000))
hs2 = ctx.hset_from_host(H[:m], 5, length=ln[:m])
for it in range(2):
    ctx.timer_start(); r = ctx.all_pairs(gm, hs2, hs2, 0.9, upper_only=True, max_out=1 << 20); ms = ctx.timer_stop()
    print("sweep %d rows: %.2f ms scored %d -> %.3e pairs/s survivors %d" % (m, ms, r["n_scored"], r["n_scored"] / ms * 1e3, r["n_out"]))
