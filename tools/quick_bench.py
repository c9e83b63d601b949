"""Quick device-timed numbers for K1 / K2 (development aid; bench.py is the contract)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from meshclust2_b200 import capi, synth

ctx = capi.Context(0)
n = int(os.environ.get("QB_N", 400000))
rng = np.random.default_rng(0)
H = rng.integers(1, 6, size=(n, 1024), dtype=np.uint8)
ln = rng.integers(950, 1050, n).astype(np.uint64)
t0 = time.time(); hs = ctx.hset_from_host(H, 5, length=ln); print("upload %.2fs" % (time.time() - t0))
for w in ("weights_cfg1_id90", "weights_appendixD_id90"):
    gm = ctx.model_from_file(os.path.join("tests", "golden", w + ".txt"))
    for flush in (True,):
        ms, nc = ctx.bench_score_pairs(gm, hs, hs, n_pairs=n, a_begin=0, b_begin=5, b_bc=1, iters=5, flush_l2=flush)
        print("%s one-vs-many n=%d: %.3f ms  %.3e pairs/s  %.1f GB/s (1057 B/pair) close=%d" % (w, n, ms, n / ms * 1e3, n * 1057 / ms / 1e6, nc))
    ia = rng.integers(0, n, n); ib = rng.integers(0, n, n)
    ms, nc = ctx.bench_score_pairs(gm, hs, hs, ia=ia, ib=ib, iters=3, flush_l2=True)
    print("%s gather pairs n=%d: %.3f ms  %.3e pairs/s  %.1f GB/s (2105 B/pair)" % (w, n, ms, n / ms * 1e3, n * 2105 / ms / 1e6))
# min-only and dot-only models
for nm, f in (("manhattan", 1 << 2), ("euclidean", 1 << 3), ("emd", 1 << 18)):
    gm = ctx.model(capi.make_desc([(f, 0.0, 1000.0)], [(0, [0])], [0.0, 1.0]))
    ms, nc = ctx.bench_score_pairs(gm, hs, hs, n_pairs=n, a_begin=0, b_begin=5, b_bc=1, iters=5, flush_l2=True)
    print("%s-only one-vs-many: %.3f ms %.3e pairs/s %.1f GB/s" % (nm, ms, n / ms * 1e3, n * 1057 / ms / 1e6))
# all-pairs sweep on 20k
m = 20000
hs2 = ctx.hset_from_host(H[:m], 5, length=ln[:m])
gm = ctx.model_from_file(os.path.join("tests", "golden", "weights_cfg1_id90.txt"))
ctx.timer_start(); r = ctx.all_pairs(gm, hs2, hs2, 0.9, upper_only=True, max_out=1 << 20); ms = ctx.timer_stop()
print("all-pairs sweep %d rows: %.1f ms, scored %d -> %.3e pairs/s, survivors %d" % (m, ms, r["n_scored"], r["n_scored"] / ms * 1e3, r["n_out"]))
# K1
seqs, _ = synth.make_set(100000, 1000, 1000, 0.08, seed=3)
t0 = time.time(); enc = capi.encode_batch(seqs); t1 = time.time()
sq = ctx.upload_seqs(enc["codes"], enc["seq_off"], enc["segs"], enc["seg_off"]); t2 = time.time()
print("encode %.3fs upload+pack %.3fs" % (t1 - t0, t2 - t1))
for k, eb in ((5, 1), (5, 2), (6, 1), (8, 2)):
    ms = ctx.bench_count_kmers(sq, k, eb, iters=3)
    print("K1 k=%d eb=%d: %.3f ms -> %.3e hist/s, %.3e kmers/s" % (k, eb, ms, 1e5 / ms * 1e3, 1e8 / ms * 1e3))
