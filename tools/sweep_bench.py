"""Device-timed all-pairs sweep on the bench's cfg3 shape (development aid; bench.py is the contract).
SB_N rows (default 100000), query rows [0, SB_Q) (default 20000), upper triangle, weights_cfg1_id90."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from meshclust2_b200 import capi, synth

n = int(os.environ.get("SB_N", 100000)); nq = int(os.environ.get("SB_Q", 20000))
ctx = capi.Context(0)
seqs, _ = synth.make_set(n, 1000, max(1, n // 100), 0.08, seed=3)
enc = capi.encode_batch(seqs)
sq = ctx.upload_seqs(enc["codes"], enc["seq_off"], enc["segs"], enc["seg_off"])
hs = ctx.count_kmers(sq, 5, 1)
for w in os.environ.get("SB_MODELS", "weights_cfg1_id90,weights_appendixD_id90").split(","):
    gm = ctx.model_from_file(os.path.join("tests", "golden", w + ".txt"))
    for rep in range(3):
        ctx.timer_start(); r = ctx.all_pairs(gm, hs, hs, 0.9, q_range=(0, nq), upper_only=True, max_out=1 << 24); ms = ctx.timer_stop()
        print("%s sweep q[0,%d) x %d: %.1f ms, scored %d -> %.3e pairs/s, survivors %d" % (w, nq, n, ms, r["n_scored"], r["n_scored"] / ms * 1e3, r["n_out"]), flush=True)
# dot-only / emd-only / min-only models
for nm, f, hi in (("euclidean", 1 << 3, 50.0), ("emd", 1 << 18, 300000.0), ("manhattan", 1 << 2, 2000.0)):
    gm = ctx.model(capi.make_desc([(f, 0.0, hi)], [(0, [0])], [-8.0, 10.0]))
    ctx.timer_start(); r = ctx.all_pairs(gm, hs, hs, 0.9, q_range=(0, nq), upper_only=True, max_out=1 << 24); ms = ctx.timer_stop()
    print("%s-only sweep: %.1f ms, scored %d -> %.3e pairs/s, survivors %d" % (nm, ms, r["n_scored"], r["n_scored"] / ms * 1e3, r["n_out"]), flush=True)
