"""Where a multi-GPU all-pairs step spends its time outside the sweep kernel (development aid).
torchrun --nproc-per-node N tools/step_phases.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as tdist
from meshclust2_b200 import capi, dist as mdist, synth

world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
tdist.init_process_group("nccl", device_id=torch.device("cuda", lr))
comm = mdist.Comm(tdist)
n_total = 100000
per, bounds = mdist.shard_bounds(n_total, world)
lo, hi = bounds[rank]
seqs, _, k, eb = synth.make_config_range("cfg3", lo, hi)
ctx = capi.Context(lr)
model = ctx.model_from_file(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "weights_cfg1_id90.txt"))
enc = capi.encode_batch(seqs)
eng = mdist.GpuEngine(capi, ctx, torch, model, k, eb, lr)
eng.set_local_sequences(ctx.upload_seqs(enc["codes"], enc["seq_off"], enc["segs"], enc["seg_off"]), hi - lo, per)
T = {}
def mark(name, t0):
    torch.cuda.synchronize(); ctx.sync()
    T[name] = T.get(name, 0.0) + (time.perf_counter() - t0)
    return time.perf_counter()
for it in range(8):
    if it == 3:
        T.clear()
    comm.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    ctx.flush_l2(256 << 20); t = mark("flush", t)
    full = getattr(eng, "_gfull", None)
    if full is None:
        eng.count_and_gather_in_place(comm, n_total); t = mark("first", t)
    else:
        eng._slot = 0
        ctx.count_kmers_into_rows(eng.seqs, full, comm.rank * eng.per); t = mark("k1", t)
        for g in eng._gt:
            comm.dist.all_gather_into_tensor(g, g[comm.rank * eng.per:(comm.rank + 1) * eng.per])
        t = mark("gathers", t)
        full.refresh(set_mag=False); t = mark("refresh", t)
    (q0, q1), = mdist.triangle_row_blocks(n_total, world, rank)
    ctx.profile(True)
    ns, sc, pairs = eng.sweep(q0, q1, True, 0.9, 1 << 22); t = mark("sweep_call", t)
    kt = ctx.kernel_time(3)[0]; ct = ctx.kernel_time(5)[0]
    ctx.profile(False)
    T["sweep_kernel"] = T.get("sweep_kernel", 0.0) + kt * 1e-3
    T["cum16_kernel"] = T.get("cum16_kernel", 0.0) + ct * 1e-3
    comm.all_reduce_sum([sc, ns], torch, eng.device); t = mark("all_reduce", t)
if rank == 0:
    print({k: round(v / 5 * 1e3, 3) for k, v in T.items()}, "ms per step (5 steps)")
ctx.close()
tdist.destroy_process_group()
