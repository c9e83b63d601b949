"""Where a distributed candidate scan's time goes (development aid): the pieces of dist.candidate_scan timed one by one with
a device synchronise after each.  python tools/scan_breakdown.py            (one rank)
torchrun --nproc-per-node 2 tools/scan_breakdown.py                          (two ranks, NCCL)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from meshclust2_b200 import capi, dist as mdist

world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
tdist = None
if world > 1:
    import torch.distributed as tdist
    tdist.init_process_group("nccl", device_id=torch.device("cuda", lr))
comm = mdist.Comm(tdist)
n_total = int(os.environ.get("SB_N", 1000000))
per, bounds = mdist.shard_bounds(n_total, world)
lo, hi = bounds[rank]
ctx = capi.Context(lr)
model = ctx.model_from_file(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "weights_cfg1_id90.txt"))
rng = np.random.default_rng(5)
base = rng.integers(1, 6, size=(10000, 1024), dtype=np.uint8)
tid = np.arange(lo, hi) % 10000
H = base[tid]
ln = (950 + (tid * 7919) % 100).astype(np.uint64)
eng = mdist.GpuEngine(capi, ctx, torch, model, 5, 1, lr)
eng.local_hset = ctx.hset_from_host(H, 5, length=ln)
eng.n_local, eng.per = hi - lo, per
del H
dev = eng.device


def T(fn, reps=50):
    torch.cuda.synchronize(); ctx.sync(); comm.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize(); ctx.sync()
    return (time.perf_counter() - t0) / reps * 1e6, r


qs = [(q * 15485863) % n_total for q in range(50)]
it = iter(qs * 100)
for q in qs[:5]:
    mdist.candidate_scan(eng, comm, torch, q, n_total, 0.9)
us, _ = T(lambda: mdist.candidate_scan(eng, comm, torch, next(it), n_total, 0.9))
out = ["whole candidate_scan %.0f us" % us]
hs, bse, n = eng._local()
us, _ = T(lambda: ctx.get_close(model, hs, bse + 5, hs, cand_begin=bse, n_cand=n, cutoff=0.9, marks=eng._marks(n)))
out.append("get_close over the shard, page-locked marks %.0f us" % us)
us, _ = T(lambda: ctx.get_close(model, hs, bse + 5, hs, cand_begin=bse, n_cand=n, cutoff=0.9))
out.append("get_close over the shard, fresh marks %.0f us" % us)
ctx.profile(True)
ctx.get_close(model, hs, bse + 5, hs, cand_begin=bse, n_cand=n, cutoff=0.9, marks=eng._marks(n))
out.append("kernel ms by kind " + str({k: ctx.kernel_time(k) for k in range(6)}))
ctx.profile(False)
us, _ = T(lambda: eng.local_query(3))
out.append("local_query %.0f us" % us)
b, l, m = eng.local_query(3)
if tdist is not None:
    pack = eng.query_pack()
    us, _ = T(lambda: tdist.broadcast(pack, src=0))
    out.append("broadcast of the query message %.0f us" % us)
us, _ = T(lambda: eng.scan_local(b, l, m, 0.9))
out.append("scan_local (stage query + get_close) %.0f us" % us)
if tdist is not None:
    def comb():
        mine = torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64, device=dev)
        allv = torch.empty((world * 3,), dtype=torch.float64, device=dev)
        tdist.all_gather_into_tensor(allv, mine)
        return allv.cpu().numpy()
    us, _ = T(comb)
    out.append("combine (tensor + all_gather + cpu) %.0f us" % us)
if rank == 0:
    print("\n".join(out), flush=True)
ctx.close()
if tdist is not None:
    tdist.destroy_process_group()
