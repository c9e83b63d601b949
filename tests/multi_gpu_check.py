#!/usr/bin/env python3
"""Multi-rank parity check on real GPUs: the row-block sharded all-pairs sweep, the range-partitioned candidate scan and the
center-sharded update stage (meshclust2_b200/dist.py) against the CPU oracle, one process per GPU over NCCL.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tests/multi_gpu_check.py [--n-seqs 3000]
  python tests/multi_gpu_check.py            # world size 1, same code path without a process group

Every rank checks its own shard of the result; exit status 0 only if all ranks agree with the oracle.
(The oracle is test infrastructure: this script is a test, not a product path.)"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-seqs", type=int, default=3000)
    ap.add_argument("--queries", type=int, default=12)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    import torch
    from meshclust2_b200 import capi, dist as mdist, synth
    from oracle import port
    tdist = None
    if world > 1:
        import torch.distributed as td
        torch.cuda.set_device(local_rank)
        td.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        tdist = td
    comm = mdist.Comm(tdist)
    n_total = args.n_seqs
    weights = os.path.join(ROOT, "tests", "golden", "weights_cfg1_id90.txt")
    k, eb, cutoff = 5, 1, 0.9
    per, bounds = mdist.shard_bounds(n_total, world)
    lo, hi = bounds[rank]
    seqs, _ = synth.make_range(n_total, 1000, 40, 0.1, seed=4242, lo=lo, hi=hi)
    ctx = capi.Context(local_rank)
    model = ctx.model_from_file(weights)
    eng = mdist.GpuEngine(capi, ctx, torch, model, k, eb, local_rank)
    enc = capi.encode_batch(seqs)
    eng.set_local_sequences(ctx.upload_seqs(enc["codes"], enc["seq_off"], enc["segs"], enc["seg_off"]), hi - lo, per)

    # oracle view of the whole set (every rank builds it: small)
    all_seqs, _ = synth.make_range(n_total, 1000, 40, 0.1, seed=4242)
    omodel = port.Model.from_text(open(weights).read())
    pts = [port.get_point(s, k, eb) for s in all_seqs]
    H = np.stack([p["hist"] for p in pts])
    mag = np.array([p["mag"] for p in pts], dtype=np.uint64)
    ln = np.array([p["len"] for p in pts], dtype=np.uint64)

    bad = 0
    # --- sweep, twice (second pass exercises the in-place buffer reuse) ---
    for rep in range(2):
        res = mdist.all_pairs_step(eng, comm, torch, n_total, cutoff, upper_only=True, blocks_per_rank=2)
        mine = sorted(map(tuple, res["survivors"].tolist()))
        want = []
        scored = 0
        for q0, q1 in res["blocks"]:
            for q in range(q0, q1):
                cand = np.arange(q + 1, n_total)
                if cand.size == 0:
                    continue
                wlo, whi = int(float(ln[q]) * cutoff), int(float(ln[q]) / cutoff)   # fastcar window, FC_Runner.cpp:427-470
                cand = cand[(ln[cand] >= wlo) & (ln[cand] <= whi)]
                if cand.size == 0:
                    continue
                r = port.score_pairs(omodel, H, mag, ln, cand.astype(np.uint64), np.full(cand.size, q, dtype=np.uint64))
                scored += int(cand.size)
                want += [(int(q), int(c)) for c in cand[r["close"].astype(bool)]]
        want.sort()
        if mine != want or res["local_scored"] != scored:
            print("[rank %d] sweep pass %d MISMATCH: %d vs %d survivors, %d vs %d scored" % (
                rank, rep, len(mine), len(want), res["local_scored"], scored), flush=True)
            bad += 1
    # --- candidate scan ---
    rng = np.random.default_rng(9)
    queries = [0, n_total - 1, per - 1 if per else 0, min(per, n_total - 1)] + rng.integers(0, n_total, args.queries).tolist()
    n_close = 0
    for q in queries:
        r = mdist.candidate_scan(eng, comm, torch, int(q), n_total, cutoff)
        ob, obd, omin, omarks = port.get_close(omodel, H, mag, ln, int(q), np.arange(n_total), cutoff)
        n_close += int(omarks.sum())
        if r["best"] != ob or (ob >= 0 and abs(r["best_dist"] - obd) > 1e-9) or r["is_min"] != omin \
                or not np.array_equal(np.asarray(r["marks_local"], dtype=np.uint8), omarks[lo:hi]):
            print("[rank %d] scan q=%d MISMATCH: best %d vs %d, dist %r vs %r, is_min %r vs %r" % (
                rank, q, r["best"], ob, r["best_dist"], obd, r["is_min"], omin), flush=True)
            bad += 1
    # --- update stage: one pass of mean_shift_update + merge over every center, centers split over ranks ---
    rng = np.random.default_rng(21)
    nc = min(400, n_total // 3)
    rows = (np.arange(nc, dtype=np.uint64) * 3) % n_total
    rows[6:9] = rows[5]
    cmag, clen = mag[rows].copy(), ln[rows].copy()
    cmag[::2] += np.uint64(23)                                   # stale magnitudes (quirk Q4)
    off, mem = [0], []
    for j, r in enumerate(rows):
        m_lo, m_hi = max(0, int(r) - 9), min(n_total, int(r) + 10 + (j % 5) * 7)
        if j % 37 == 2:
            m_hi = m_lo
        mem.append(np.arange(m_lo, m_hi, dtype=np.uint64))
        off.append(off[-1] + m_hi - m_lo)
    off, members, delta = np.array(off, dtype=np.uint64), np.concatenate(mem), 4
    nxt, ng = mdist.update_pass(eng, comm, torch, rows, cmag, clen, off, members, cutoff)
    mg = mdist.merge_pass(eng, comm, torch, rows, cmag, clen, delta, cutoff)
    H2 = np.vstack([H, H[rows.astype(np.int64)]])
    mag2, ln2 = np.concatenate([mag, cmag]).astype(np.uint64), np.concatenate([ln, clen]).astype(np.uint64)
    Hc, idx = H[rows.astype(np.int64)], np.arange(nc)
    n_moved = n_merged = 0
    for c in range(rank, nc, world):                             # every rank checks a stride of the (replicated) result
        memc = members[int(off[c]):int(off[c + 1])]
        want_next, want_good = -1, 0
        if len(memc):
            keep = port.filter_members(omodel, H2, mag2, ln2, n_total + c, memc, cutoff).astype(bool)
            want_good = int(keep.sum())
            if want_good:
                want_next = int(np.flatnonzero(keep)[port.mean_closest(H, memc[keep])[0]])
        last = min(nc - 1, c + delta)
        want_merge = port.merge(omodel, Hc, cmag, clen, idx, c, c + 1, last, cutoff) if last >= c + 1 else 0
        n_moved += want_next >= 0
        n_merged += want_merge > c
        if nxt[c] != want_next or ng[c] != want_good or mg[c] != want_merge:
            print("[rank %d] update stage center %d MISMATCH: next %d vs %d, good %d vs %d, merge %d vs %d" % (
                rank, c, nxt[c], want_next, ng[c], want_good, mg[c], want_merge), flush=True)
            bad += 1
    # every rank looked at a stride of the centers: the check is vacuous only if NO rank saw a move or a merge
    tot_moved, tot_merged = comm.all_reduce_sum([int(n_moved), int(n_merged)], torch, eng.device)
    if tot_moved == 0 or tot_merged == 0:
        if rank == 0:
            print("[rank 0] update stage check is vacuous (moved %d, merged %d over all ranks)" % (tot_moved, tot_merged), flush=True)
        bad += 1
    tot_bad = comm.all_reduce_sum([bad], torch, eng.device)[0]
    if rank == 0:
        print("multi_gpu_check world=%d n=%d: %s (sweep scored %d close %d; %d scan queries, %d marks)" % (
            world, n_total, "OK" if tot_bad == 0 else "FAILED", res["n_scored"], res["n_close"], len(queries), n_close), flush=True)
    if tdist is not None:
        tdist.destroy_process_group()
    sys.exit(0 if tot_bad == 0 else 1)


if __name__ == "__main__":
    main()
