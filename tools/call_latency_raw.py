"""True per-call latency of the small-batch C-ABI entry points: ctypes arguments are built once, the loop only calls."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from meshclust2_b200 import capi
ctx = capi.Context(0)
L = capi.lib()
rng = np.random.default_rng(0)
n = 20000
H = rng.integers(1, 7, size=(n, 1024), dtype=np.uint8)
ln = rng.integers(950, 1050, n).astype(np.uint64)
hs = ctx.hset_from_host(H, 5, length=ln)
sc = ctx.hset_from_host(H[:64], 5, length=ln[:64])
gm = ctx.model_from_file(os.path.join("tests", "golden", "weights_cfg1_id90.txt"))
def loop(fn, reps=2000):
    for _ in range(50):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e6
for m in (8, 64, 512, 4096):
    cand = rng.integers(0, n, m).astype(np.uint64)
    marks = np.zeros(m, dtype=np.uint8)
    best, bd, ismin = C.c_int64(), C.c_double(), C.c_int32()
    pc, pm = cand.ctypes.data_as(C.c_void_p), marks.ctypes.data_as(C.c_void_p)
    args = (ctx.h, gm.h, sc.h, C.c_uint64(0), hs.h, pc, C.c_uint64(0), C.c_uint64(m), C.c_double(0.9), C.byref(best), C.byref(bd), C.byref(ismin), pm)
    t_gc = loop(lambda: L.mc2_get_close(*args))
    fargs = (ctx.h, gm.h, sc.h, C.c_uint64(0), hs.h, pc, C.c_uint64(m), C.c_double(0.9), pm)
    t_fl = loop(lambda: L.mc2_filter(*fargs))
    mean = np.ones(1024); pmean = mean.ctypes.data_as(C.c_void_p)
    cargs = (ctx.h, hs.h, pc, C.c_uint64(m), pmean, C.byref(best), C.byref(bd), None)
    t_cl = loop(lambda: L.mc2_closest(*cargs))
    print("m=%5d  get_close %6.1f us  filter %6.1f us  closest %6.1f us" % (m, t_gc, t_fl, t_cl))
d = np.array([0], dtype=np.uint64); s = np.array([5], dtype=np.uint64); mg = np.array([1100], dtype=np.uint64); le = np.array([1000], dtype=np.uint64)
aargs = (ctx.h, sc.h, C.c_uint64(1), d.ctypes.data_as(C.c_void_p), hs.h, s.ctypes.data_as(C.c_void_p), mg.ctypes.data_as(C.c_void_p), le.ctypes.data_as(C.c_void_p))
print("assign_rows(1) %.1f us (stream-ordered, no sync)" % loop(lambda: L.mc2_hset_assign_rows(*aargs)))
print("assign_rows(1) + get_close(64): %.1f us" % loop(lambda: (L.mc2_hset_assign_rows(*aargs), L.mc2_get_close(*args))))
print("ctx sync only: %.1f us" % loop(lambda: L.mc2_ctx_sync(ctx.h)))
