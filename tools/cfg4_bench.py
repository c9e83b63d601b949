"""BASELINE configs[3] shape: long --single-file records (5 contigs x 10 kb joined by 50 N), k=8, uint16 histograms
(65,536 bins = 128 KiB per row): K1 rate and the all-pairs sweep rate (HBM-bound form: 131,105 B per pair)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from meshclust2_b200 import capi, synth
n = int(os.environ.get("CFG4_N", 5000))
ctx = capi.Context(0)
t0 = time.time()
seqs = synth.make_single_file(n, 5, 10000, seed=4)
enc = capi.encode_batch(seqs)
print("generated %d records, %.1f kb each, in %.1fs" % (n, np.mean([len(s) for s in seqs]) / 1e3, time.time() - t0))
sq = ctx.upload_seqs(enc["codes"], enc["seq_off"], enc["segs"], enc["seg_off"])
L = np.mean([len(s) for s in seqs])
for rep in range(2):
    ms = ctx.bench_count_kmers(sq, 8, 2, iters=3, flush_l2=True)
byts = n * (L / 4 + 65536 * 2 + 40)
print("K1 k=8 u16: %.2f ms  %.3e hist/s  %.0f GB/s (%.1f%% of 6540)" % (ms, n / ms * 1e3, byts / ms / 1e6, byts / ms / 1e6 / 65.4))
hs, largest, eb = ctx.count_kmers_auto(sq, 8)
print("auto width: largest count %d -> %d bytes" % (largest, eb))
if eb != 2:
    hs = ctx.count_kmers(sq, 8, 2)
gm = ctx.model_from_file(os.path.join("tests", "golden", os.environ.get("PT_W", "weights_cfg1_id90") + ".txt"))
for rep in range(int(os.environ.get('CFG4_REPS', 2))):
    ctx.timer_start(); r = ctx.all_pairs(gm, hs, hs, 0.9, upper_only=True, max_out=1 << 22); ms = ctx.timer_stop()
    print("sweep %d x %d upper: %.1f ms scored %d -> %.3e pairs/s = %.0f GB/s at 131105 B/pair (%.1f%% of 6540), survivors %d" % (
        n, n, ms, r["n_scored"], r["n_scored"] / ms * 1e3, r["n_scored"] * 131105 / ms / 1e6, r["n_scored"] * 131105 / ms / 1e6 / 65.4, r["n_out"]))
# candidate form: one query vs all
for rep in range(2):
    ms, nc = ctx.bench_score_pairs(gm, hs, hs, n_pairs=n, a_begin=0, b_begin=5, b_bc=1, iters=3, flush_l2=True)
print("one-vs-many %d: %.2f ms %.3e pairs/s %.0f GB/s (%.1f%% of 6540)" % (n, ms, n / ms * 1e3, n * 131105 / ms / 1e6, n * 131105 / ms / 1e6 / 65.4))
