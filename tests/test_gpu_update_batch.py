"""GPU: the batched update / merge stage (mc2_update_centers, mc2_merge_centers) vs the oracle and vs the one-center calls.

mean_shift_update (src/cluster/ClusterFactory.cpp:288-335) = Trainer::filter + per-bin mean of the survivors +
Trainer::closest, and Trainer::merge (src/cluster/Trainer.cpp:74-109), for every center of one pass in one launch.  The
results are positions / indices, so the comparison is exact."""
import numpy as np
import pytest

from oracle import port
from meshclust2_b200 import synth
from conftest import weights_path, weights_text

pytestmark = pytest.mark.gpu


def _points(capi, ctx, n, k, eb, seed):
    seqs, _ = synth.make_set(n, 1000, 10, 0.07, seed=seed)
    enc = capi.encode_batch(seqs, threads=4)
    sq = ctx.upload_seqs(enc["codes"], enc["seq_off"], enc["segs"], enc["seg_off"])
    hs = ctx.count_kmers(sq, k, eb)
    got = hs.download()
    return hs, got["hist"], got["mag"].astype(np.uint64), got["len"].astype(np.uint64)


def _centers(rng, n, mag, ln, n_centers):
    rows = rng.integers(0, n, n_centers)
    cmag = mag[rows].copy()
    clen = ln[rows].copy()
    stale = rng.random(n_centers) < 0.5            # quirk Q4: a center keeps the magnitude it was constructed with
    cmag[stale] += rng.integers(1, 60, stale.sum()).astype(np.uint64)
    clen[rng.random(n_centers) < 0.2] -= np.uint64(7)
    return rows, cmag, clen


def _stage(ctx, hs, k, eb, rows, cmag, clen):
    nc = len(rows)
    sc = ctx.hset_from_host(np.ones((nc, 4 ** k), dtype=port.DTYPES[eb]), k, length=np.ones(nc, dtype=np.uint64))
    sc.assign_rows(np.arange(nc), hs, rows, mag=cmag, length=clen)
    return sc


@pytest.mark.parametrize("k,eb,seed", [(5, 1, 1), (4, 2, 2), (5, 1, 3)])
def test_update_centers_vs_oracle_and_single_calls(built_lib, ctx, k, eb, seed):
    rng = np.random.default_rng(seed)
    n, nc = 260, 48
    hs, H, mag, ln = _points(built_lib, ctx, n, k, eb, seed)
    rows, cmag, clen = _centers(rng, n, mag, ln, nc)
    sc = _stage(ctx, hs, k, eb, rows, cmag, clen)
    m = port.Model.from_text(weights_text("weights_cfg1_id90"))
    gm = ctx.model_from_file(weights_path("weights_cfg1_id90"))
    sizes = rng.integers(0, 70, nc)
    sizes[[3, 17]] = 0                              # empty member lists
    sizes[5] = 300                                  # more members than one CTA has threads
    off = np.zeros(nc + 1, dtype=np.uint64)
    off[1:] = np.cumsum(sizes)
    members = rng.integers(0, n, int(off[-1])).astype(np.uint64)
    nxt, ng = ctx.update_centers(gm, sc, nc, hs, off, members, 0.9)
    # oracle: the centers as extra rows (bins of the point they carry, their own magnitude / length)
    H2 = np.vstack([H, H[rows]])
    mag2 = np.concatenate([mag, cmag]).astype(np.uint64)
    ln2 = np.concatenate([ln, clen]).astype(np.uint64)
    some_survive = False
    for c in range(nc):
        mem = members[int(off[c]):int(off[c + 1])]
        if len(mem) == 0:
            assert nxt[c] == -1 and ng[c] == 0
            continue
        keep = port.filter_members(m, H2, mag2, ln2, n + c, mem, 0.9).astype(bool)
        assert ng[c] == keep.sum(), c
        assert np.array_equal(keep, ctx.filter(gm, sc, c, hs, mem, 0.9).astype(bool))
        if not keep.any():
            assert nxt[c] == -1
            continue
        some_survive = True
        pos = np.flatnonzero(keep)
        ob = port.mean_closest(H, mem[keep])[0]
        assert nxt[c] == pos[ob], (c, nxt[c], pos[ob])
        gb = ctx.mean_closest(hs, mem[keep])[0]
        assert gb == ob
    assert some_survive


def test_update_centers_edge_cases(built_lib, ctx):
    rng = np.random.default_rng(9)
    hs, H, mag, ln = _points(built_lib, ctx, 64, 5, 1, 5)
    gm = ctx.model_from_file(weights_path("weights_cfg1_id90"))
    rows = np.array([7, 7, 12])                     # two centers carrying the same point
    sc = _stage(ctx, hs, 5, 1, rows, mag[rows], ln[rows])
    # every list empty
    nxt, ng = ctx.update_centers(gm, sc, 3, hs, np.zeros(4, dtype=np.uint64), np.zeros(0, dtype=np.uint64), 0.9)
    assert (nxt == -1).all() and (ng == 0).all()
    # a center whose only member is the point it carries: it survives and is chosen
    nxt, ng = ctx.update_centers(gm, sc, 3, hs, [0, 1, 2, 3], [7, 7, 12], 0.9)
    assert nxt.tolist() == [0, 0, 0] and ng.tolist() == [1, 1, 1]
    # duplicates: the first of equal distances wins (strict < in Trainer::closest)
    nxt, ng = ctx.update_centers(gm, sc, 1, hs, [0, 4], [7, 7, 7, 7], 0.9)
    assert nxt[0] == 0 and ng[0] == 4
    # members outside the length window are dropped before the mean
    ln2 = ln.copy(); ln2[20] = 4000
    hs2 = ctx.hset_from_host(H, 5, mag=mag, length=ln2)
    nxt, ng = ctx.update_centers(gm, sc, 1, hs2, [0, 2], [20, 7], 0.9)
    assert nxt[0] == 1 and ng[0] == 1
    with pytest.raises(built_lib.Mc2Error):
        ctx.update_centers(gm, sc, 1, hs, [0, 1], [64], 0.9)          # member row out of range


@pytest.mark.parametrize("delta", [0, 1, 5, 100])
def test_merge_centers_vs_oracle_and_single_calls(built_lib, ctx, delta):
    rng = np.random.default_rng(40 + delta)
    n, nc = 200, 60
    hs, H, mag, ln = _points(built_lib, ctx, n, 5, 1, 11)
    rows, cmag, clen = _centers(rng, n, mag, ln, nc)
    rows[10:14] = rows[9]                           # runs of near-identical centers so that merges happen
    cmag[10:14] = cmag[9]; clen[10:14] = clen[9]
    sc = _stage(ctx, hs, 5, 1, rows, cmag, clen)
    m = port.Model.from_text(weights_text("weights_cfg1_id90"))
    gm = ctx.model_from_file(weights_path("weights_cfg1_id90"))
    out = ctx.merge_centers(gm, sc, nc, delta, 0.9)
    Hc, idx = H[rows], np.arange(nc)
    merged = 0
    for c in range(nc):
        last = min(nc - 1, c + delta)
        want = port.merge(m, Hc, cmag, clen, idx, c, c + 1, last, 0.9) if last >= c + 1 else 0
        assert out[c] == want, (c, out[c], want)
        assert out[c] == ctx.merge(gm, sc, idx, c, c + 1, last, 0.9)
        merged += out[c] > c
    assert delta == 0 or merged > 0
