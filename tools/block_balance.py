"""Time the folded query-row blocks of the cfg3 sweep one by one on one GPU: how even is the per-rank work at N ranks?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from meshclust2_b200 import capi, dist as mdist, synth
N = int(os.environ.get("BB_WORLD", 8)); bpr = int(os.environ.get("BB_BPR", 2))
ctx = capi.Context(0)
seqs, _, k, eb = synth.make_config_range("cfg3", 0, 100000)
hs = ctx.count_kmers(ctx.seqs_from_text(seqs), k, eb)
gm = ctx.model_from_file(os.path.join("tests", "golden", "weights_cfg1_id90.txt"))
n = len(seqs)
tot = []
for rank in range(N):
    blocks = mdist.folded_row_blocks(n, N, rank, bpr)
    ms_r = []
    for q0, q1 in blocks:
        for rep in range(2):
            ctx.timer_start(); r = ctx.all_pairs(gm, hs, hs, 0.9, q_range=(q0, q1), upper_only=True, max_out=1 << 22); ms = ctx.timer_stop()
        ms_r.append((q0, q1, ms, r["n_scored"]))
    tot.append(sum(m[2] for m in ms_r))
    print("rank %d: %s -> %.2f ms" % (rank, ", ".join("[%d,%d) %.2f ms %.2e pairs" % m for m in ms_r), tot[-1]))
print("max %.2f  mean %.2f  imbalance %.1f%%" % (max(tot), np.mean(tot), 100 * (max(tot) / np.mean(tot) - 1)))
