"""GPU: the tile form of the all-pairs sweep (csrc/tile_sweep.cu: TMA ring, tcgen05 u8 Gram term in TMEM, u16 cumulative
EMD, fp32 screen + exact fp64 epilogue) through the C ABI.
Integer reductions bit-exact against numpy restatements of Feature.cpp:858-871 (manhattan), :1112-1124 (dot term of
euclidean) and :1504-1518 (emd); survivors, scores (1e-9) and counts against the oracle's sweep."""
import numpy as np
import pytest

from conftest import weights_path, weights_text
from helpers import FAST, all_singles_model, synth_hist, to_desc
from oracle import port

pytestmark = pytest.mark.gpu


def _np_reductions(Hq, Hd):
    q = Hq.astype(np.int64)
    d = Hd.astype(np.int64)
    dot = q @ d.T
    cq, cd = np.cumsum(q, axis=1), np.cumsum(d, axis=1)
    sad = np.zeros((len(q), len(d)), dtype=np.int64)
    emd = np.zeros_like(sad)
    for i in range(len(q)):
        sad[i] = np.abs(d - q[i]).sum(axis=1)
        emd[i] = np.abs(cd - cq[i]).sum(axis=1)
    return dot, emd, sad


def _hist(rng, n, hi):
    """uint8 rows whose sums stay below 65536; hi steers the row sum (which selects the packed-sum flush interval)"""
    return rng.integers(0, hi + 1, size=(n, 1024)).astype(np.uint8)


@pytest.mark.parametrize("hi,need", [(3, 7), (3, 6), (3, 4), (3, 2), (3, 1), (3, 5), (3, 3), (20, 7), (60, 7)])
def test_tile_reductions_bit_exact(built_lib, ctx, hi, need):
    """row sums ~1.5 k (flush every 16 words), ~10 k (every 4), ~31 k (dp2a per word); ragged tiles on both sides"""
    rng = np.random.default_rng(100 + hi + need)
    nq, nd = 150, 300
    Hq, Hd = _hist(rng, nq, hi), _hist(rng, nd, hi)
    Hd[7] = 0  # an all-zero row and a saturated-ish row
    Hq[3] = min(hi * 2, 63)
    hq = ctx.hset_from_host(Hq, 5, mag=None, length=np.full(nq, 1000, dtype=np.uint64))
    hd = ctx.hset_from_host(Hd, 5, mag=None, length=np.full(nd, 1000, dtype=np.uint64))
    for (q0, q1), (d0, d1) in [((0, nq), (0, nd)), ((5, 133), (130, 290)), ((64, 65), (0, 1)), ((0, 64), (128, 256))]:
        got = ctx.tile_reductions(hq, hd, need, (q0, q1), (d0, d1))
        dot, emd, sad = _np_reductions(Hq[q0:q1], Hd[d0:d1])
        if need & 2:
            assert np.array_equal(got["dot"].astype(np.int64), dot), "sum p*q (tcgen05 u8 MMA) differs"
        if need & 4:
            assert np.array_equal(got["emd"].astype(np.int64), emd), "sum |cumP-cumQ| differs"
        if need & 1:
            assert np.array_equal(got["sad"].astype(np.int64), sad), "sum |p-q| differs"


def test_tile_reductions_same_set_many_tiles(built_lib, ctx):
    """one set against itself over several tiles of the schedule (more tiles than one CTA sees once)"""
    rng = np.random.default_rng(5)
    n = 700
    H = synth_hist(rng, n, 5, 1, hi=9)
    hs = ctx.hset_from_host(H, 5, mag=None, length=np.full(n, 1000, dtype=np.uint64))
    got = ctx.tile_reductions(hs, hs, 7)
    dot, emd, sad = _np_reductions(H, H)
    assert np.array_equal(got["dot"].astype(np.int64), dot)
    assert np.array_equal(got["emd"].astype(np.int64), emd)
    assert np.array_equal(got["sad"].astype(np.int64), sad)
    assert (np.diag(got["emd"]) == 0).all() and (np.diag(got["sad"]) == 0).all()


def _oracle_sweep(om, H, mag, ln, cutoff, q_range, d_range, upper):
    """the reference's work() loop (FC_Runner.cpp:427-470) through the oracle: survivors {(q, d): score}, #scored"""
    out, scored = {}, 0
    for q in range(*q_range):
        lo, hi = int(float(ln[q]) * cutoff), int(float(ln[q]) / cutoff)
        cand = [d for d in range(*d_range) if lo <= int(ln[d]) <= hi and (not upper or d > q)]
        if not cand:
            continue
        cand = np.array(cand)
        r = port.score_pairs(om, H, mag, ln, cand, np.full(len(cand), q))
        scored += len(cand)
        for d, s, c in zip(cand, r["score"], r["close"]):
            if c:
                out[(q, int(d))] = s
    return out, scored


@pytest.mark.parametrize("wname", ["weights_cfg1_id90", "weights_appendixD_id90"])
def test_tile_sweep_vs_oracle_trained_models(built_lib, ctx, wname):
    rng = np.random.default_rng(11)
    n = 520
    H = synth_hist(rng, n, 5, 1, hi=6)
    ln = rng.integers(850, 1150, n).astype(np.uint64)
    mag = H.sum(axis=1, dtype=np.uint64)
    hs = ctx.hset_from_host(H, 5, mag=None, length=ln)
    gm = ctx.model_from_file(weights_path(wname))
    om = port.Model.from_text(weights_text(wname))
    for (qr, dr, upper) in [((0, n), (0, n), True), ((100, 230), (0, n), False), ((3, 70), (60, 200), True)]:
        r = ctx.all_pairs(gm, hs, hs, 0.9, q_range=qr, d_range=dr, upper_only=upper, max_out=n * n)
        want, scored = _oracle_sweep(om, H, mag, ln, 0.9, qr, dr, upper)
        got = {(int(q), int(d)): s for q, d, s in zip(r["q"], r["d"], r["score"])}
        assert r["n_scored"] == scored
        edge = {k for k, s in want.items() if abs(s - 0.5) <= 1e-9}
        assert set(got) - edge == set(want) - edge
        for k in set(got) & set(want):
            assert abs(got[k] - want[k]) <= 1e-9
        assert r["n_out"] == len(got)
        assert len(want) > 0


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_tile_sweep_random_models(built_lib, ctx, seed):
    """random models over the nine fast singles (every reduction mix the kernel is instantiated for): the fp32 screen
    must never drop a pair the exact epilogue calls close"""
    from meshclust2_b200 import capi
    rng = np.random.default_rng(seed)
    n = 260
    H = synth_hist(rng, n, 5, 1, hi=5)
    ln = rng.integers(900, 1100, n).astype(np.uint64)
    mag = H.sum(axis=1, dtype=np.uint64)
    if seed % 2:
        mag = mag.copy()
        mag[::3] += 13  # stale pseudo-magnitudes (quirk Q4)
    flags = list(rng.choice(FAST, size=int(rng.integers(1, 6)), replace=False))
    om = all_singles_model([int(f) for f in flags], H, mag, ln, rng, n_combos=int(rng.integers(1, 5)))
    om.weights[0] = float(om.weights[0]) - 0.5
    hs = ctx.hset_from_host(H, 5, mag=mag, length=ln)
    gm = ctx.model(to_desc(capi, om))
    r = ctx.all_pairs(gm, hs, hs, 0.9, upper_only=True, max_out=n * n)
    want, scored = _oracle_sweep(om, H, mag, ln, 0.9, (0, n), (0, n), True)
    got = {(int(q), int(d)): s for q, d, s in zip(r["q"], r["d"], r["score"])}
    assert r["n_scored"] == scored
    edge = {k for k, s in want.items() if abs(s - 0.5) <= 1e-9}
    assert set(got) - edge == set(want) - edge
    for k in set(got) & set(want):
        assert abs(got[k] - want[k]) <= 1e-9


def _wide_hist(rng, n, k, eb, mean_extra):
    """rows of 4^k bins >= 1 (the pseudo-count) plus sparse k-mer counts: about mean_extra counted k-mers per row"""
    N = 4 ** k
    H = np.ones((n, N), dtype=np.int64)
    for r in range(n):
        idx = rng.integers(0, N, mean_extra)
        np.add.at(H[r], idx, 1)
    return H.astype(port.DTYPES[eb])


@pytest.mark.parametrize("k,eb,extra", [(6, 1, 30000), (6, 2, 30000), (6, 2, 64000), (7, 2, 50000), (8, 2, 50000), (5, 2, 3000)])
def test_tile_reductions_wide_rows_bit_exact(built_lib, ctx, k, eb, extra):
    """rows of several 1 KiB slabs (k = 6 .. 8) and uint16 bins: row sums reach 65536 only through the 4^k pseudo-counts, which
    the cumulative rows leave out (the differences cumP - cumQ do not change); uint16 bins travel as a byte plane"""
    rng = np.random.default_rng(1000 * k + eb)
    nq, nd = (70, 150) if k < 8 else (66, 130)
    Hq, Hd = _wide_hist(rng, nq, k, eb, extra), _wide_hist(rng, nd, k, eb, extra)
    Hd[5] = 1                                               # nothing but pseudo-counts
    hq = ctx.hset_from_host(Hq, k, mag=None, length=np.full(nq, 1000, dtype=np.uint64))
    hd = ctx.hset_from_host(Hd, k, mag=None, length=np.full(nd, 1000, dtype=np.uint64))
    got = ctx.tile_reductions(hq, hd, 7)
    dot, emd, sad = _np_reductions(Hq, Hd)
    assert np.array_equal(got["dot"].astype(np.int64), dot)
    assert np.array_equal(got["emd"].astype(np.int64), emd)
    assert np.array_equal(got["sad"].astype(np.int64), sad)


@pytest.mark.parametrize("k,eb", [(6, 1), (6, 2), (7, 2)])
def test_tile_sweep_wide_rows_vs_oracle(built_lib, ctx, k, eb):
    """the sweep over multi-slab rows through the tile form == the oracle's work() loop (every in-window pair takes the exact
    epilogue there: the fp32 screen is for 1 KiB rows)"""
    from meshclust2_b200 import capi
    rng = np.random.default_rng(77 + k + eb)
    n = 200
    base = _wide_hist(rng, 25, k, eb, 20000).astype(np.int64)
    H = base[rng.integers(0, 25, n)]
    for r in range(n):                                      # related rows: a few hundred extra k-mers each
        np.add.at(H[r], rng.integers(0, 4 ** k, 300), 1)
    H = H.astype(port.DTYPES[eb])
    ln = rng.integers(18000, 22000, n).astype(np.uint64)
    mag = H.sum(axis=1, dtype=np.uint64)
    flags = [port.FEAT[x] for x in ("euclidean", "emd", "manhattan", "pearson", "length_difference")]
    om = all_singles_model(flags, H, mag, ln, rng, n_combos=4)
    # intercept such that about one pair in ten is close
    iu, ju = np.triu_indices(n, 1)
    sc = np.clip(port.score_pairs(om, H, mag, ln, ju, iu)["score"], 1e-12, 1 - 1e-12)
    om.weights[0] = float(om.weights[0]) - float(np.quantile(np.log(sc / (1 - sc)), 0.9))
    hs = ctx.hset_from_host(H, k, mag=mag, length=ln)
    gm = ctx.model(to_desc(capi, om))
    r = ctx.all_pairs(gm, hs, hs, 0.9, upper_only=True, max_out=n * n)
    want, scored = _oracle_sweep(om, H, mag, ln, 0.9, (0, n), (0, n), True)
    got = {(int(q), int(d)): s for q, d, s in zip(r["q"], r["d"], r["score"])}
    assert r["n_scored"] == scored
    edge = {kk for kk, s in want.items() if abs(s - 0.5) <= 1e-9}
    assert set(got) - edge == set(want) - edge
    for kk in set(got) & set(want):
        assert abs(got[kk] - want[kk]) <= 1e-9
    assert 0 < len(want) < scored
    # the same through the row-streaming kernels
    import os
    os.environ["MC2_SWEEP_LEGACY"] = "1"
    try:
        r2 = ctx.all_pairs(gm, hs, hs, 0.9, upper_only=True, max_out=n * n)
    finally:
        os.environ.pop("MC2_SWEEP_LEGACY", None)
    got2 = {(int(q), int(d)): s for q, d, s in zip(r2["q"], r2["d"], r2["score"])}
    assert got2 == got and r2["n_scored"] == r["n_scored"]


def test_tile_sweep_declines_unfit_rows(built_lib, ctx):
    """a uint16 bin above 255, or a zero bin in rows whose sums need the pseudo-counts taken out: the tile form declines
    (nothing written) and the sweep is served by the row-streaming kernels with the same answers as the oracle"""
    from meshclust2_b200 import capi
    rng = np.random.default_rng(5)
    n, k = 120, 6
    for case in ("big_bin", "zero_bin"):
        H = _wide_hist(rng, n, k, 2, 30000 if case == "big_bin" else 64000).astype(np.int64)
        if case == "big_bin":
            H[7, 100] = 300
        else:
            H[9, 5] = 0
        H = H.astype(np.uint16)
        ln = rng.integers(19000, 21000, n).astype(np.uint64)
        mag = H.sum(axis=1, dtype=np.uint64)
        hs = ctx.hset_from_host(H, k, mag=mag, length=ln)
        with pytest.raises(capi.Mc2Error):
            ctx.tile_reductions(hs, hs, 7)
        om = all_singles_model([port.FEAT["euclidean"], port.FEAT["emd"]], H, mag, ln, rng, n_combos=2)
        gm = ctx.model(to_desc(capi, om))
        r = ctx.all_pairs(gm, hs, hs, 0.9, upper_only=True, max_out=n * n)
        want, scored = _oracle_sweep(om, H, mag, ln, 0.9, (0, n), (0, n), True)
        got = {(int(q), int(d)): s for q, d, s in zip(r["q"], r["d"], r["score"])}
        assert r["n_scored"] == scored
        edge = {kk for kk, s in want.items() if abs(s - 0.5) <= 1e-9}
        assert set(got) - edge == set(want) - edge


@pytest.mark.parametrize("bias", [0.3, -0.3, 0.6, -0.7])
def test_tile_sweep_with_a_bias(built_lib, ctx, bias):
    """--bias moves the decision to logistic(sum) >= 0.5 - bias (Predictor.cpp:323-333): the screen's threshold follows it;
    at |bias| >= 0.5 every pair (or none) is close and all of them take the exact epilogue"""
    from meshclust2_b200 import capi
    rng = np.random.default_rng(3)
    n = 300
    H = synth_hist(rng, n, 5, 1, hi=6)
    ln = rng.integers(900, 1100, n).astype(np.uint64)
    mag = H.sum(axis=1, dtype=np.uint64)
    om = port.Model.from_text(weights_text("weights_cfg1_id90"))
    om.bias = bias
    hs = ctx.hset_from_host(H, 5, mag=None, length=ln)
    gm = ctx.model(to_desc(capi, om))
    r = ctx.all_pairs(gm, hs, hs, 0.9, upper_only=True, max_out=n * n)
    want, scored = _oracle_sweep(om, H, mag, ln, 0.9, (0, n), (0, n), True)
    got = {(int(q), int(d)): s for q, d, s in zip(r["q"], r["d"], r["score"])}
    assert r["n_scored"] == scored
    edge = {k for k, s in want.items() if abs(s - 0.5) <= 1e-9}
    assert set(got) - edge == set(want) - edge
    for k in set(got) & set(want):
        assert abs(got[k] - want[k]) <= 1e-9
    if bias >= 0.5:
        assert len(want) == scored
    if bias <= -0.5:
        assert len(want) == 0
