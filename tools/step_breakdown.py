"""Wall-clock breakdown of one bench step at N=1 (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from meshclust2_b200 import capi, dist as mdist, synth
n = int(os.environ.get("SB_N", 100000))
seqs, _, k, eb = synth.make_config_range("cfg3", 0, n, n=n)
enc = capi.encode_batch(seqs)
ctx = capi.Context(0)
model = ctx.model_from_file("tests/golden/weights_cfg1_id90.txt")
eng = mdist.GpuEngine(capi, ctx, torch, model, k, eb, 0)
eng.set_local_sequences(ctx.upload_seqs(enc["codes"], enc["seq_off"], enc["segs"], enc["seg_off"]), n, n)
comm = mdist.Comm(None)
def T():
    ctx.sync(); return time.perf_counter()
for rep in range(3):
    t0 = T(); ctx.flush_l2(256 << 20); t1 = T()
    eng.count(); t2 = T()
    eng.use_local_as_full(); t3 = T()
    blocks = mdist.folded_row_blocks(n, 1, 0, 8)
    tk = 0.0; tw = 0.0
    for q0, q1 in blocks:
        a = T(); r = eng.sweep(q0, q1, True, 0.9, 1 << 22); b = T(); tw += b - a
    t4 = T()
    comm.all_reduce_sum([1, 2], torch, eng.device); t5 = T()
    print("rep %d: flush %.1f ms, count %.1f ms, use_local %.1f ms, sweeps %.1f ms (wall sum %.1f), allreduce %.1f ms, total %.1f ms" % (
        rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, tw * 1e3, (t5 - t4) * 1e3, (t5 - t0) * 1e3))
ctx.profile(True)
for q0, q1 in blocks:
    eng.sweep(q0, q1, True, 0.9, 1 << 22)
print("kernel time sweep:", ctx.kernel_time(3))
