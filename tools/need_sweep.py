"""One-vs-many (2^20 candidates, 1 GiB) timing per NEED mask: which reductions cost what."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from meshclust2_b200 import capi
ctx = capi.Context(0)
rng = np.random.default_rng(0)
n = 1 << 20
base = rng.integers(1, 7, size=(4096, 1024), dtype=np.uint8)
H = base[rng.integers(0, 4096, n)]
ln = rng.integers(950, 1050, n).astype(np.uint64)
hs = ctx.hset_from_host(H, 5, length=ln)
F = dict(lengthd=1 << 21, manhattan=1 << 2, euclidean=1 << 3, emd=1 << 18)
def model(names):
    return ctx.model(capi.make_desc([(F[x], 0.0, 1000.0) for x in names], [(0, [i]) for i in range(len(names))], [0.0] + [1.0] * len(names)))
for names in (["lengthd"], ["manhattan"], ["euclidean"], ["emd"], ["manhattan", "euclidean"], ["euclidean", "emd"], ["manhattan", "euclidean", "emd"]):
    gm = model(names)
    for rep in range(2):
        ms, nc = ctx.bench_score_pairs(gm, hs, hs, n_pairs=n, a_begin=0, b_begin=5, b_bc=1, iters=10, flush_l2=False)
    print("%-32s %.3f ms  %.3e pairs/s  %.0f GB/s  (%.1f%% of 6540)" % ("+".join(names), ms, n / ms * 1e3, n * 1057 / ms / 1e6, n * 1057 / ms / 1e6 / 65.4))
# gather list (both rows streamed) for the full model
gm = model(["manhattan", "euclidean", "emd"])
ia = rng.integers(0, n, n); ib = rng.integers(0, n, n)
ms, nc = ctx.bench_score_pairs(gm, hs, hs, ia=ia, ib=ib, iters=5, flush_l2=False)
print("gather pairs (2 rows/pair): %.3f ms %.3e pairs/s %.0f GB/s (2105 B/pair)" % (ms, n / ms * 1e3, n * 2105 / ms / 1e6))
