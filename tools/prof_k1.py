"""Small fixed workload for ncu captures of K1: 100k x 1 kb reads, k=5, uint8 (cfg3 shape), three counting launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshclust2_b200 import capi, synth
ctx = capi.Context(0)
seqs, _, k, eb = synth.make_config_range("cfg3", 0, int(os.environ.get("K1_N", 100000)))
enc = capi.encode_batch(seqs)
sq = ctx.upload_seqs(enc["codes"], enc["seq_off"], enc["segs"], enc["seg_off"])
ms = ctx.bench_count_kmers(sq, int(os.environ.get("K1_K", 5)), int(os.environ.get("K1_EB", 1)), iters=3, flush_l2=True)
print("K1 %.3f ms" % ms)
