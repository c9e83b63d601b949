"""Per-call latency of the small-batch entry points the mean-shift driver issues one at a time (get_close / filter / merge /
assign_rows / closest) - wall clock per call including the ctypes hop (~3 us)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from meshclust2_b200 import capi
ctx = capi.Context(0)
rng = np.random.default_rng(0)
n = 20000
H = rng.integers(1, 7, size=(n, 1024), dtype=np.uint8)
ln = rng.integers(950, 1050, n).astype(np.uint64)
hs = ctx.hset_from_host(H, 5, length=ln)
sc = ctx.hset_from_host(H[:64], 5, length=ln[:64])
gm = ctx.model_from_file(os.path.join("tests", "golden", "weights_cfg1_id90.txt"))
def timeit(f, reps=300):
    for _ in range(20):
        f()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    return (time.perf_counter() - t0) / reps * 1e6
for m in (8, 64, 512, 4096, 16384):
    cand = rng.integers(0, n, m).astype(np.uint64)
    t_gc = timeit(lambda: ctx.get_close(gm, sc, 0, hs, cand=cand, cutoff=0.9))
    t_fl = timeit(lambda: ctx.filter(gm, sc, 0, hs, cand, 0.9))
    t_cl = timeit(lambda: ctx.closest(hs, cand, np.ones(1024)))
    print("m=%6d  get_close %7.1f us   filter %7.1f us   closest %7.1f us" % (m, t_gc, t_fl, t_cl))
t_as = timeit(lambda: sc.assign_rows([0], hs, [5], mag=[1100], length=[1000]))
rows = np.arange(11, dtype=np.uint64)
t_mg = timeit(lambda: ctx.merge(gm, sc, rows, 0, 1, 10, 0.9))
print("assign_rows(1) %.1f us   merge(10) %.1f us" % (t_as, t_mg))
