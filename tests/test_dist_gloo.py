"""CPU, world_size 2, gloo: the N>1 host logic of meshclust2_b200/dist.py (contiguous K1 shards, padded all-gather of
histogram shards, folded row-block split of the upper-triangular sweep).  The compute engine here is a stand-in built
on the oracle (test infrastructure); on GPUs the same driver runs GpuEngine over the C ABI."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from conftest import weights_text
from meshclust2_b200 import dist as mdist
from meshclust2_b200 import synth
from oracle import port


def test_shard_bounds_and_blocks_cover_everything():
    for n in (1, 7, 64, 1000, 1001):
        for world in (1, 2, 3, 8):
            per, b = mdist.shard_bounds(n, world)
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert all(hi - lo <= per for lo, hi in b)
            rows = []
            for r in range(world):
                for q0, q1 in mdist.folded_row_blocks(n, world, r):
                    rows.extend(range(q0, q1))
            assert sorted(rows) == list(range(n)), (n, world)
            rows = [q for r in range(world) for q0, q1 in mdist.rect_row_blocks(n, world, r) for q in range(q0, q1)]
            assert sorted(rows) == list(range(n))
            for bpr in (1, 3):
                rows = [q for r in range(world) for q0, q1 in mdist.triangle_row_blocks(n, world, r, bpr) for q in range(q0, q1)]
                assert rows == list(range(n)), (n, world, bpr)


def test_folded_blocks_balance_the_triangle():
    n, world = 100000, 8
    work = []
    for r in range(world):
        work.append(sum((n - 1 - q0 + n - q1) * (q1 - q0) / 2 for q0, q1 in mdist.folded_row_blocks(n, world, r)))
    assert max(work) / min(work) < 1.01


def test_triangle_ranges_balance_the_triangle():
    """one contiguous range per rank, equal pair counts (the ranges grow towards the end of the set)"""
    n, world = 100000, 8
    work = []
    for r in range(world):
        (q0, q1), = mdist.triangle_row_blocks(n, world, r)
        work.append((n - 1 - q0 + n - q1) * (q1 - q0) / 2)
    assert sum(work) == n * (n - 1) / 2
    assert max(work) / min(work) < 1.01          # equal up to the per-row tile overhead the split also counts


class OracleEngine:
    def __init__(self, seqs, k, eb, model, per):
        self.seqs, self.k, self.eb, self.model, self.per = seqs, k, eb, model, per
        self.device = torch.device("cpu")
        self.N = 4 ** k

    def count(self):
        pts = [port.get_point(s, self.k, self.eb) for s in self.seqs]
        self.H = np.stack([p["hist"] for p in pts]) if pts else np.zeros((0, self.N), dtype=port.DTYPES[self.eb])
        self.ln = np.array([p["len"] for p in pts], dtype=np.int64)
        self.mag = np.array([p["mag"] for p in pts], dtype=np.int64)

    def use_local_as_full(self):
        self.full = (self.H, self.ln, self.mag)

    def export_local(self):
        bins = torch.ones((self.per, self.N), dtype=torch.uint8)
        ln = torch.zeros((self.per,), dtype=torch.int64)
        mag = torch.zeros((self.per,), dtype=torch.int64)
        n = len(self.seqs)
        bins[:n] = torch.from_numpy(self.H)
        ln[:n] = torch.from_numpy(self.ln)
        mag[:n] = torch.from_numpy(self.mag)
        return bins, ln, mag

    def install_full(self, bins, length, mag, n_total):
        assert bins.shape[0] == n_total
        self.full = (bins.numpy(), length.numpy(), mag.numpy())

    def empty_query(self):
        return (torch.ones((1, self.N), dtype=torch.uint8), torch.zeros((1,), dtype=torch.int64),
                torch.zeros((1,), dtype=torch.int64))

    def local_query(self, row_local):
        return (torch.from_numpy(self.H[row_local:row_local + 1].copy()), torch.from_numpy(self.ln[row_local:row_local + 1].copy()),
                torch.from_numpy(self.mag[row_local:row_local + 1].copy()))

    def scan_local(self, bins, length, mag, cutoff):
        n = self.H.shape[0]
        if n == 0:
            return -1, -1.0, True, np.zeros(0, dtype=np.uint8)
        H = np.concatenate([self.H, bins.numpy()])
        ln = np.concatenate([self.ln, length.numpy()]).astype(np.uint64)
        mg = np.concatenate([self.mag, mag.numpy()]).astype(np.uint64)
        return port.get_close(self.model, H, mg, ln, n, np.arange(n), cutoff)

    def update_centers(self, rows, mag, length, off, members, cutoff):
        H, ln, mg = self.full
        H2 = np.concatenate([H, H[np.asarray(rows, dtype=np.int64)]])
        mag2 = np.concatenate([mg, mag]).astype(np.uint64)
        ln2 = np.concatenate([ln, length]).astype(np.uint64)
        nxt, ng = np.full(len(rows), -1, dtype=np.int64), np.zeros(len(rows), dtype=np.int64)
        for c in range(len(rows)):
            mem = np.asarray(members[int(off[c]):int(off[c + 1])], dtype=np.uint64)
            if len(mem) == 0:
                continue
            keep = port.filter_members(self.model, H2, mag2, ln2, H.shape[0] + c, mem, cutoff).astype(bool)
            ng[c] = keep.sum()
            if keep.any():
                nxt[c] = np.flatnonzero(keep)[port.mean_closest(H, mem[keep])[0]]
        return nxt, ng

    def merge_centers(self, rows, mag, length, delta, cutoff):
        H, _, _ = self.full
        Hc = H[np.asarray(rows, dtype=np.int64)]
        n = len(rows)
        idx = np.arange(n)
        out = np.zeros(n, dtype=np.int64)
        for c in range(n):
            last = min(n - 1, c + delta)
            if last >= c + 1:
                out[c] = port.merge(self.model, Hc, np.asarray(mag, dtype=np.uint64), np.asarray(length, dtype=np.uint64), idx,
                                    c, c + 1, last, cutoff)
        return out

    def sweep(self, q0, q1, upper_only, cutoff, max_out):
        H, ln, mag = self.full
        n = H.shape[0]
        ia, ib = [], []
        for q in range(q0, q1):
            lo, hi = int(float(ln[q]) * cutoff), int(float(ln[q]) / cutoff)
            for c in range(q + 1 if upper_only else 0, n):
                if lo <= ln[c] <= hi:
                    ia.append(c), ib.append(q)
        if not ia:
            return 0, 0, np.zeros((0, 2), dtype=np.uint64)
        o = port.score_pairs(self.model, H, mag.astype(np.uint64), ln.astype(np.uint64), ia, ib, want_cache=False)
        cl = o["close"].astype(bool)
        pairs = np.stack([np.array(ib)[cl], np.array(ia)[cl]], axis=1).astype(np.uint64)
        return int(cl.sum()), len(ia), pairs


def _scan_worker(rank, world, port_no, n_total, queries, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = port.Model.from_text(weights_text("weights_cfg1_id90"))
        per, bounds = mdist.shard_bounds(n_total, world)
        lo, hi = bounds[rank]
        seqs, _ = synth.make_range(n_total, 1000, 6, 0.1, seed=77, lo=lo, hi=hi)
        eng = OracleEngine(seqs, 5, 1, model, per)
        eng.count()
        comm = mdist.Comm(tdist)
        res = []
        for q in queries:
            r = mdist.candidate_scan(eng, comm, torch, q, n_total, 0.9)
            marks = [None] * world
            tdist.all_gather_object(marks, r["marks_local"].tolist())
            res.append((r["best"], round(r["best_dist"], 12), r["is_min"], sum(marks, [])))
        if rank == 0:
            out.put(res)
    finally:
        tdist.destroy_process_group()


def _worker(rank, world, port_no, n_total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = port.Model.from_text(weights_text("weights_cfg1_id90"))
        per, bounds = mdist.shard_bounds(n_total, world)
        lo, hi = bounds[rank]
        seqs, _ = synth.make_range(n_total, 1000, 9, 0.1, seed=31, lo=lo, hi=hi)
        eng = OracleEngine(seqs, 5, 1, model, per)
        comm = mdist.Comm(tdist)
        res = mdist.all_pairs_step(eng, comm, torch, n_total, 0.9, upper_only=True, blocks_per_rank=2)
        gathered = [None] * world
        tdist.all_gather_object(gathered, res["survivors"].tolist())
        if rank == 0:
            out.put((res["n_scored"], res["n_close"], sorted(map(tuple, sum(gathered, [])))))
    finally:
        tdist.destroy_process_group()


def _update_case(n_total, seed=5):
    """centers = every 3rd point (some with a stale magnitude), members = a window of points around each"""
    rng = np.random.default_rng(seed)
    seqs, _ = synth.make_range(n_total, 1000, 5, 0.08, seed=123)
    pts = [port.get_point(s, 5, 1) for s in seqs]
    H = np.stack([p["hist"] for p in pts])
    mag = np.array([p["mag"] for p in pts], dtype=np.uint64)
    ln = np.array([p["len"] for p in pts], dtype=np.uint64)
    rows = np.arange(0, n_total, 3, dtype=np.uint64)
    rows[6:9] = rows[5]                                      # a run of identical centers across the rank boundary: merges happen
    rows[10:12] = rows[9]
    cmag, clen = mag[rows].copy(), ln[rows].copy()
    cmag[::2] += np.uint64(17)
    off, mem = [0], []
    for j, r in enumerate(rows):
        lo, hi = max(0, int(r) - 7), min(n_total, int(r) + 8 + (j % 4) * 5)
        if j == 2:
            hi = lo                                          # an empty member list
        mem.append(np.arange(lo, hi, dtype=np.uint64))
        off.append(off[-1] + hi - lo)
    return seqs, (H, ln.astype(np.int64), mag.astype(np.int64)), rows, cmag, clen, np.array(off, dtype=np.uint64), np.concatenate(mem)


def _update_worker(rank, world, port_no, n_total, delta, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = port.Model.from_text(weights_text("weights_cfg1_id90"))
        seqs, full, rows, cmag, clen, off, members = _update_case(n_total)
        eng = OracleEngine([], 5, 1, model, 0)
        eng.full = full                                      # the replicated point set (after the all-gather)
        comm = mdist.Comm(tdist)
        nxt, ng = mdist.update_pass(eng, comm, torch, rows, cmag, clen, off, members, 0.9)
        mg = mdist.merge_pass(eng, comm, torch, rows, cmag, clen, delta, 0.9)
        if rank == 0:
            out.put((nxt.tolist(), ng.tolist(), mg.tolist()))
    finally:
        tdist.destroy_process_group()


def test_balanced_ranges_cover_and_balance():
    for world in (1, 2, 3, 8):
        for w in ([], [5], [0, 0, 0], list(range(100)), [1000] + [1] * 50):
            r = mdist.balanced_ranges(w, world)
            assert len(r) == world and r[0][0] == 0 and r[-1][1] == len(w)
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1)) and all(lo <= hi for lo, hi in r)
    r = mdist.balanced_ranges(np.ones(1000) * 50, 8)
    sizes = [hi - lo for lo, hi in r]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_sharded_update_and_merge_pass_equal_single_process():
    """mean_shift_update / merge for every center, centers split over 2 ranks == the same pass in one process"""
    n_total, delta = 60, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    pno = _free_port()
    procs = [ctx.Process(target=_update_worker, args=(r, 2, pno, n_total, delta, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    model = port.Model.from_text(weights_text("weights_cfg1_id90"))
    seqs, full, rows, cmag, clen, off, members = _update_case(n_total)
    eng = OracleEngine([], 5, 1, model, 0)
    eng.full = full
    nxt, ng = eng.update_centers(rows, cmag, clen, off, members, 0.9)
    mg = eng.merge_centers(rows, cmag, clen, delta, 0.9)
    assert got == (nxt.tolist(), ng.tolist(), mg.tolist())
    assert (nxt >= 0).any() and nxt[2] == -1 and (mg > 0).any()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(300)
def test_two_ranks_equal_one_rank():
    n_total = 61      # odd on purpose: the tail shard is padded for the all-gather
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, _free_port_cached(), n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single process, same data, no process group
    model = port.Model.from_text(weights_text("weights_cfg1_id90"))
    seqs, _ = synth.make_range(n_total, 1000, 9, 0.1, seed=31)
    eng = OracleEngine(seqs, 5, 1, model, n_total)
    res = mdist.all_pairs_step(eng, mdist.Comm(None), torch, n_total, 0.9, upper_only=True, blocks_per_rank=2)
    want = (res["n_scored"], res["n_close"], sorted(map(tuple, res["survivors"].tolist())))
    assert got == want and want[1] > 0


_PORT = []


def _free_port_cached():
    if not _PORT:
        _PORT.append(_free_port())
    return _PORT[0]


@pytest.mark.timeout(300)
def test_sharded_candidate_scan_equals_single_process():
    """get_close with the candidates range-partitioned over 2 ranks == the oracle's get_close over the whole set"""
    n_total, queries = 37, [0, 5, 18, 19, 36]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    pno = _free_port()
    procs = [ctx.Process(target=_scan_worker, args=(r, 2, pno, n_total, queries, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    model = port.Model.from_text(weights_text("weights_cfg1_id90"))
    seqs, _ = synth.make_range(n_total, 1000, 6, 0.1, seed=77)
    pts = [port.get_point(s, 5, 1) for s in seqs]
    H = np.stack([p["hist"] for p in pts])
    mag = np.array([p["mag"] for p in pts], dtype=np.uint64)
    ln = np.array([p["len"] for p in pts], dtype=np.uint64)
    for (best, bd, ismin, marks), qq in zip(got, queries):
        ob = port.get_close(model, H, mag, ln, qq, np.arange(n_total), 0.9)
        assert best == ob[0] and abs(bd - ob[1]) <= 1e-9 and ismin == ob[2] and marks == ob[3].tolist()
