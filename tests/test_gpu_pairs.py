"""GPU: K2 (pair features + GLM + cutoff) through the C ABI vs the oracle and the reference-generated golden vectors.
Integer-valued singles bit-exact; floating-point singles, caches and scores within REL_TOL = 1e-9 (north_star: 1e-6);
close flags identical except pairs whose score sits within 1e-9 of the 0.5 decision boundary."""
import numpy as np
import pytest

from conftest import weights_path, weights_text
from helpers import FAST, SLOW, REL_TOL, all_singles_model, assert_close_rel, assert_flags_match, synth_hist, to_desc
from oracle import port

pytestmark = pytest.mark.gpu

INT_SINGLES = {port.FEAT["manhattan"], port.FEAT["emd"], port.FEAT["length_difference"]}


def _side(H, rng, stale=False):
    n = H.shape[0]
    mag = H.sum(axis=1, dtype=np.uint64)
    if stale:
        mag = mag.copy()
        mag[::3] += 13
        mag[1::5] -= 3
    ln = rng.integers(800, 1300, n).astype(np.uint64)
    return mag, ln


@pytest.mark.parametrize("wname", ["weights_cfg1_id90", "weights_appendixD_id90"])
def test_golden_classifier(built_lib, ctx, golden, wname):
    H, ln, mag = golden["hist_k5_eb1"], golden["len_k5_eb1"], golden["mag_k5_eb1"]
    hs = ctx.hset_from_host(H, 5, mag=None, length=ln)
    gm = ctx.model_from_file(weights_path(wname))
    ja, jb = golden["score_ia"], golden["score_ib"]
    g = ctx.score_pairs(gm, hs, hs, ja, jb)
    assert np.abs(g["score"] - golden[wname + "_score"]).max() <= 1e-9
    assert_close_rel(g["cache"], golden[wname + "_cache"], REL_TOL, "cache")
    assert_close_rel(g["dist"], golden[wname + "_dist"], 1e-8, "dist")
    assert_flags_match(g["close"], golden[wname + "_close"], golden[wname + "_score"])
    assert g["close"].any()
    # stale pseudo-magnitudes (quirk Q4): host-supplied mag is what intersection/kulczynski2/pearson must read
    hs2 = ctx.hset_from_host(H, 5, mag=golden["stale_mag_k5_eb1"], length=ln)
    g2 = ctx.score_pairs(gm, hs2, hs2, ja, jb, want=("score", "close"))
    assert np.abs(g2["score"] - golden[wname + "_stale_score"]).max() <= 1e-9
    assert_flags_match(g2["close"], golden[wname + "_stale_close"], golden[wname + "_stale_score"])
    # callers
    cand = golden["cand"]
    for t, q in enumerate(golden[wname + "_gc_q"]):
        best, bd, ismin, marks = ctx.get_close(gm, hs, int(q), hs, cand=cand, cutoff=0.9)
        assert best == golden[wname + "_gc_best"][t]
        assert abs(bd - golden[wname + "_gc_dist"][t]) <= 1e-9
        assert ismin == bool(golden[wname + "_gc_ismin"][t])
        assert np.array_equal(marks, golden[wname + "_gc_marks"][t])
        keep = ctx.filter(gm, hs, int(q), hs, cand, 0.9)
        assert np.array_equal(keep, golden[wname + "_filter_keep"][t])
        rows = cand[(np.arange(8) + int(q)) % len(cand)]
        assert ctx.merge(gm, hs, rows, 0, 1, 7, 0.9) == golden[wname + "_merge"][t]


@pytest.mark.parametrize("k,eb", [(5, 1), (3, 2), (2, 4), (4, 8)])
def test_golden_raw_singles(built_lib, ctx, golden, k, eb):
    """every in-scope raw single for every width against the reference's own static Feature<T>::xxx outputs"""
    tag = "k%d_eb%d" % (k, eb)
    H, ln = golden["hist_" + tag], golden["len_" + tag]
    names = [str(x) for x in golden["single_names"]]
    ia, ib = golden["pair_ia"], golden["pair_ib"]
    # zero lengths make length_difference throw and all-ones histograms make pearson NaN (then the reference throws
    # in normalize_cache); both are tested separately as error cases
    ok = np.isfinite(golden["raw_" + tag]).all(axis=1)
    ia, ib, want = ia[ok], ib[ok], golden["raw_" + tag][ok]
    assert ok.sum() > 100
    hs = ctx.hset_from_host(H, k, length=ln)
    flags = [port.FEAT[n] for n in names]
    singles = [(f, 0.0, 1.0) for f in flags]
    combos = [(0, [i]) for i in range(len(flags))]
    gm = ctx.model(built_lib.make_desc(singles, combos, [0.0] * (len(flags) + 1)))
    g = ctx.score_pairs(gm, hs, hs, ia, ib, want=("raw",))
    for c, nm in enumerate(names):
        fin = np.isfinite(want[:, c])
        if nm in ("manhattan", "emd", "length_difference"):
            assert np.array_equal(g["raw"][fin, c], want[fin, c]), nm
        else:
            assert_close_rel(g["raw"][fin, c], want[fin, c], REL_TOL, tag + " " + nm)


def test_golden_raw_singles_stale(built_lib, ctx, golden):
    H, ln, mag = golden["hist_k5_eb1"], golden["len_k5_eb1"], golden["stale_mag_k5_eb1"]
    names = [str(x) for x in golden["single_names"]]
    ia, ib = golden["pair_ia"], golden["pair_ib"]
    ok = np.isfinite(golden["raw_stale_k5_eb1"]).all(axis=1)
    ia, ib, want = ia[ok], ib[ok], golden["raw_stale_k5_eb1"][ok]
    hs = ctx.hset_from_host(H, 5, mag=mag, length=ln)
    flags = [port.FEAT[n] for n in names]
    gm = ctx.model(built_lib.make_desc([(f, 0.0, 1.0) for f in flags], [(0, [i]) for i in range(len(flags))],
                                       [0.0] * (len(flags) + 1)))
    g = ctx.score_pairs(gm, hs, hs, ia, ib, want=("raw",))
    fin = np.isfinite(want)
    assert_close_rel(np.where(fin, g["raw"], 0), np.where(fin, want, 0), REL_TOL, "stale raw")


@pytest.mark.parametrize("k,eb,flags", [
    (5, 1, FAST), (6, 1, FAST), (4, 1, FAST), (2, 1, FAST), (1, 1, FAST),     # fast u8 path (k>=5) and generic (small k)
    (5, 2, FAST), (4, 2, FAST), (3, 2, FAST), (7, 2, FAST),                  # fast u16 path (k>=4) and generic
    (3, 4, FAST), (5, 4, FAST), (3, 8, FAST), (4, 8, FAST),                  # wide types: reference's wrap-around arithmetic
    (5, 1, SLOW), (4, 2, SLOW), (3, 4, SLOW), (2, 8, SLOW),                  # + jefferey / jensen-shannon
])
def test_random_models_vs_oracle(built_lib, ctx, k, eb, flags):
    rng = np.random.default_rng(1000 * k + 10 * eb + len(flags))
    n = 96
    hi = {1: 255, 2: 3000, 4: 100000, 8: 100000}[eb]
    H = synth_hist(rng, n, k, eb, hi=hi if k <= 5 else min(hi, 40))
    mag, ln = _side(H, rng, stale=True)
    model = all_singles_model(flags, H, mag, ln, rng)
    hs = ctx.hset_from_host(H, k, mag=mag, length=ln)
    gm = ctx.model(to_desc(built_lib, model))
    m = 700
    ia, ib = rng.integers(0, n, m), rng.integers(0, n, m)
    g = ctx.score_pairs(gm, hs, hs, ia, ib)
    o = port.score_pairs(model, H, mag, ln, ia, ib)
    what = "k=%d eb=%d S=%d" % (k, eb, len(flags))
    assert_close_rel(g["cache"], o["cache"], 1e-8, what + " cache")
    assert np.abs(g["score"] - o["score"]).max() <= 1e-8, what
    assert_flags_match(g["close"], o["close"], o["score"], tol=1e-8, what=what)
    assert not g["skipped"].any()
    # the straight-line epilogue (taken when neither cache nor raw is requested; constant divisions through a reciprocal
    # with an exact correction step, 64-bit Pearson) must give the very same bits as the interpretive one above
    f = ctx.score_pairs(gm, hs, hs, ia, ib, want=("score", "dist", "close"))
    assert np.array_equal(f["score"], g["score"]) and np.array_equal(f["dist"], g["dist"]), what
    assert np.array_equal(f["close"], g["close"]), what
    # lazy form (no score requested: the logistic is evaluated only where the decision needs it)
    z = ctx.score_pairs(gm, hs, hs, ia, ib, want=("dist", "close"))
    assert np.array_equal(z["close"], g["close"]) and np.array_equal(z["dist"], g["dist"]), what


@pytest.mark.parametrize("eb", [1, 2])
def test_extreme_histograms(built_lib, ctx, eb):
    """all-ones, saturated, one-hot rows: exact integer reductions must survive the extremes"""
    k, N = 5, 1024
    dt = port.DTYPES[eb]
    # uint16: the reference squares T-promoted ints, so bins > 46340 overflow `int` there (undefined behaviour);
    # the CUDA path returns the exact value instead, so the extreme case stops just below that line
    top = int(np.iinfo(dt).max) if eb == 1 else 46000
    H = np.ones((6, N), dtype=dt)
    H[1, :] = top                       # saturated everywhere
    H[2, 0] = top                       # one-hot first bin
    H[3, N - 1] = top                   # one-hot last bin (largest EMD)
    H[4, ::2] = top
    H[5, :] = np.arange(N) % 200 + 1
    ln = np.full(6, 1000, dtype=np.uint64)
    names = ["manhattan", "euclidean", "normalized_vectors", "intersection", "emd", "kulczynski2", "simratio"]
    flags = [port.FEAT[n] for n in names]
    hs = ctx.hset_from_host(H, k, length=ln)
    gm = ctx.model(built_lib.make_desc([(f, 0.0, 1.0) for f in flags], [(0, [i]) for i in range(len(flags))],
                                       [0.0] * (len(flags) + 1)))
    ia, ib = np.repeat(np.arange(6), 6), np.tile(np.arange(6), 6)
    g = ctx.score_pairs(gm, hs, hs, ia, ib, want=("raw",))
    mag = H.sum(axis=1, dtype=np.uint64)
    for j, (a, b) in enumerate(zip(ia, ib)):
        for c, nm in enumerate(names):
            want = port.raw_single(port.FEAT[nm], H[a], H[b], int(mag[a]), int(mag[b]), 1000, 1000)
            got = g["raw"][j, c]
            if nm in ("manhattan", "emd"):
                assert got == want, (nm, a, b, got, want)
            else:
                assert abs(got - want) <= REL_TOL * max(abs(want), 1e-300) or (np.isnan(got) and np.isnan(want)), (nm, a, b)


def test_length_filter_and_broadcast_forms(built_lib, ctx, golden):
    H, ln, mag = golden["hist_k5_eb1"], golden["len_k5_eb1"], golden["mag_k5_eb1"]
    cand = golden["cand"]
    Hc, lnc, magc = H[cand], ln[cand].copy(), mag[cand]
    lnc[::4] = (lnc[::4] * 0.7).astype(np.uint64)          # push a quarter of the lengths out of the 0.9 window
    lnc[1::6] = (lnc[1::6] * 1.3).astype(np.uint64)
    hs = ctx.hset_from_host(Hc, 5, length=lnc)
    m = port.Model.from_text(weights_text("weights_cfg1_id90"))
    gm = ctx.model_from_file(weights_path("weights_cfg1_id90"))
    n = len(cand)
    for q in (0, 5, 11):
        # contiguous candidate range (no index list) and explicit list must agree with the oracle's get_close
        for kw in (dict(cand=None, cand_begin=0, n_cand=n), dict(cand=np.arange(n))):
            best, bd, ismin, marks = ctx.get_close(gm, hs, q, hs, cutoff=0.9, **kw)
            ob = port.get_close(m, Hc, magc, lnc, q, np.arange(n), 0.9)
            assert best == ob[0] and abs(bd - ob[1]) <= 1e-9 and ismin == ob[2] and np.array_equal(marks, ob[3])
        r = ctx.score_pairs(gm, hs, hs, ia=np.arange(n), b_begin=q, b_bc=1, len_filter=1, anchor_is_b=1, cutoff=0.9)
        lo, hi = int(float(lnc[q]) * 0.9), int(float(lnc[q]) / 0.9)
        want_skip = (lnc < lo) | (lnc > hi)
        assert np.array_equal(r["skipped"].astype(bool), want_skip) and want_skip.any() and not want_skip.all()
        assert np.isnan(r["score"][want_skip]).all() and not r["close"][want_skip].any()
        keep = ctx.filter(gm, hs, q, hs, np.arange(n), 0.9)
        assert np.array_equal(keep, port.filter_members(m, Hc, magc, lnc, q, np.arange(n), 0.9))


def test_get_close_none_in_window_and_empty(built_lib, ctx, golden):
    H, ln = golden["hist_k5_eb1"][:8], golden["len_k5_eb1"][:8].copy()
    ln[0] = 5000
    hs = ctx.hset_from_host(H, 5, length=ln)
    gm = ctx.model_from_file(weights_path("weights_cfg1_id90"))
    best, bd, ismin, marks = ctx.get_close(gm, hs, 0, hs, cand=np.arange(1, 8), cutoff=0.9)
    assert best == -1 and bd == -1 and ismin and not marks.any()
    best, bd, ismin, marks = ctx.get_close(gm, hs, 0, hs, cand=np.zeros(0, dtype=np.uint64), cutoff=0.9)
    assert best == -1 and ismin
    assert ctx.merge(gm, hs, np.arange(8), 0, 1, 7, 0.9) == 0


def test_zero_length_and_nan_are_errors(built_lib, ctx, golden):
    H, ln = golden["hist_k5_eb1"][:4], golden["len_k5_eb1"][:4].copy()
    gm = ctx.model_from_file(weights_path("weights_cfg1_id90"))       # uses length_difference and pearson
    ln0 = ln.copy(); ln0[1] = 0
    hs = ctx.hset_from_host(H, 5, length=ln0)
    with pytest.raises(built_lib.Mc2Error) as e:                      # Feature::length_difference throws 123
        ctx.score_pairs(gm, hs, hs, [0], [1])
    assert e.value.status == -4
    Hc = H.copy(); Hc[2, :] = 1                                        # zero variance -> pearson NaN -> reference throws
    hs = ctx.hset_from_host(Hc, 5, length=ln)
    with pytest.raises(built_lib.Mc2Error) as e:
        ctx.score_pairs(gm, hs, hs, [0], [2])
    assert e.value.status == -4


def test_unsupported_feature_flag_is_rejected(built_lib, ctx):
    d = built_lib.make_desc([(1 << 1, 0.0, 1.0)], [(0, [0])], [0.0, 1.0])      # FEAT_HELLINGER: out of scope (a9)
    with pytest.raises(built_lib.Mc2Error) as e:
        ctx.model(d)
    assert e.value.status == -5


def test_regression_model_clamps(built_lib, ctx, golden):
    """Predictor::p_predict (similarity): sum clamped to [0,1], no logistic"""
    H, ln, mag = golden["hist_k5_eb1"], golden["len_k5_eb1"], golden["mag_k5_eb1"]
    m = port.Model.from_text(weights_text("weights_appendixD_id90"))
    hs = ctx.hset_from_host(H, 5, length=ln)
    gm = ctx.model(to_desc(built_lib, m, regression=1))
    ja, jb = golden["score_ia"][:100], golden["score_ib"][:100]
    g = ctx.score_pairs(gm, hs, hs, ja, jb, want=("score",))
    want = np.array([port.predict_pair(m, H[a], H[b], int(mag[a]), int(mag[b]), int(ln[a]), int(ln[b])) for a, b in zip(ja, jb)])
    assert np.abs(g["score"] - want).max() <= 1e-9 and (want == 0).any() and (want > 0).any()


def test_bias_shifts_decisions(built_lib, ctx, golden):
    H, ln, mag = golden["hist_k5_eb1"], golden["len_k5_eb1"], golden["mag_k5_eb1"]
    m = port.Model.from_text(weights_text("weights_cfg1_id90"))
    m.bias = 0.3
    hs = ctx.hset_from_host(H, 5, length=ln)
    gm = ctx.model(to_desc(built_lib, m))
    ja, jb = golden["score_ia"], golden["score_ib"]
    g = ctx.score_pairs(gm, hs, hs, ja, jb, want=("score", "close"))
    o = port.score_pairs(m, H, mag, ln, ja, jb)
    assert np.abs(g["score"] - o["score"]).max() <= 1e-9 and np.array_equal(g["close"], o["close"])
    assert g["close"].sum() > golden["weights_cfg1_id90_close"].sum()


def test_distance(built_lib, ctx, golden):
    H = golden["hist_k5_eb1"]
    hs = ctx.hset_from_host(H, 5, length=golden["len_k5_eb1"])
    got = ctx.distance(hs, hs, golden["pair_ia"], golden["pair_ib"])
    assert np.array_equal(got, golden["distance_k5_eb1"])


def test_symmetry_and_batch_equals_single(built_lib, ctx, golden):
    H, ln = golden["hist_k5_eb1"], golden["len_k5_eb1"]
    hs = ctx.hset_from_host(H, 5, length=ln)
    gm = ctx.model_from_file(weights_path("weights_appendixD_id90"))
    ja, jb = golden["score_ia"][:64], golden["score_ib"][:64]
    ab = ctx.score_pairs(gm, hs, hs, ja, jb, want=("score", "raw"))
    ba = ctx.score_pairs(gm, hs, hs, jb, ja, want=("score", "raw"))
    assert_close_rel(ab["raw"], ba["raw"], 1e-12, "symmetric singles")
    one = np.array([ctx.score_pairs(gm, hs, hs, [a], [b], want=("score",))["score"][0] for a, b in zip(ja, jb)])
    assert np.array_equal(one, ab["score"])          # batched == one-by-one, bitwise


def test_set_row_keeps_stale_mag(built_lib, ctx, golden):
    """DivergencePoint::set copies bins + length but not mag (quirk Q4): mc2_hset_set_row mirrors that"""
    H, ln = golden["hist_k5_eb1"][:6], golden["len_k5_eb1"][:6]
    hs = ctx.hset_from_host(H, 5, length=ln)
    centers = ctx.hset_from_host(H[:2], 5, length=ln[:2])
    centers.set_row(0, hs, 4)
    got = centers.download()
    assert np.array_equal(got["hist"][0], H[4]) and got["len"][0] == ln[4]
    assert got["mag"][0] == H[0].sum()                # still the old point's magnitude
    m = port.Model.from_text(weights_text("weights_appendixD_id90"))
    gm = ctx.model_from_file(weights_path("weights_appendixD_id90"))
    g = ctx.score_pairs(gm, centers, hs, [0], [3], want=("score",))
    Hc, magc = np.stack([H[4], H[3]]), np.array([H[0].sum(), H[3].sum()], dtype=np.uint64)
    o = port.score_pairs(m, Hc, magc, np.array([ln[4], ln[3]]), [0], [1])
    assert abs(g["score"][0] - o["score"][0]) <= 1e-9


def test_all_pairs_sweep_vs_oracle(built_lib, ctx, golden):
    """fastcar-style query-vs-database sweep with the length window and fused cutoff (FC_Runner.cpp:427-470)"""
    H, ln, mag = golden["hist_k5_eb1"], golden["len_k5_eb1"], golden["mag_k5_eb1"]
    cand = golden["cand"]
    Hc, lnc, magc = H[cand], ln[cand], mag[cand]
    n = len(cand)
    hs = ctx.hset_from_host(Hc, 5, length=lnc)
    m = port.Model.from_text(weights_text("weights_cfg1_id90"))
    gm = ctx.model_from_file(weights_path("weights_cfg1_id90"))
    r = ctx.all_pairs(gm, hs, hs, 0.9, upper_only=True, max_out=n * n)
    want = set()
    scored = 0
    ia, ib = [], []
    for q in range(n):
        lo, hi = int(float(lnc[q]) * 0.9), int(float(lnc[q]) / 0.9)
        for c in range(q + 1, n):
            if lo <= lnc[c] <= hi:
                ia.append(c), ib.append(q)
    o = port.score_pairs(m, Hc, magc, lnc, ia, ib)          # close(pts[i], query): database row first
    want = {(b, a) for a, b, cl in zip(ia, ib, o["close"]) if cl}
    got = set(zip(r["q"].tolist(), r["d"].tolist()))
    assert r["n_scored"] == len(ia) and got == want and len(want) > 0
    # capacity smaller than the survivor count: count still exact, payload truncated
    r2 = ctx.all_pairs(gm, hs, hs, 0.9, upper_only=True, max_out=3)
    assert r2["n_out"] == len(want) and len(r2["q"]) == 3
    # rectangular block, not upper-only
    r3 = ctx.all_pairs(gm, hs, hs, 0.9, q_range=(2, 9), d_range=(0, n), upper_only=False, max_out=n * n)
    want3 = set()
    for q in range(2, 9):
        lo, hi = int(float(lnc[q]) * 0.9), int(float(lnc[q]) / 0.9)
        cs = [c for c in range(n) if lo <= lnc[c] <= hi]
        oo = port.score_pairs(m, Hc, magc, lnc, cs, [q] * len(cs))
        want3 |= {(q, c) for c, cl in zip(cs, oo["close"]) if cl}
    assert set(zip(r3["q"].tolist(), r3["d"].tolist())) == want3


def test_full_size_properties_100k_one_vs_many(built_lib, ctx):
    """BASELINE-size candidate scan (1 query x 100k candidates, k=5, u8): batch result equals chunked results bitwise,
    scores symmetric under swapping sides, and a strided sample agrees with the oracle."""
    rng = np.random.default_rng(9)
    n = 100000
    H = synth_hist(rng, n, 5, 1, hi=40)
    ln = rng.integers(950, 1050, n).astype(np.uint64)
    hs = ctx.hset_from_host(H, 5, length=ln)
    m = port.Model.from_text(weights_text("weights_cfg1_id90"))
    gm = ctx.model_from_file(weights_path("weights_cfg1_id90"))
    full = ctx.score_pairs(gm, hs, hs, a_begin=0, n_pairs=n, b_begin=7, b_bc=1, want=("score", "close"))
    parts = [ctx.score_pairs(gm, hs, hs, a_begin=s, n_pairs=25000, b_begin=7, b_bc=1, want=("score",))["score"]
             for s in range(0, n, 25000)]
    assert np.array_equal(np.concatenate(parts), full["score"])
    idx = np.arange(0, n, 997)
    mag = H.sum(axis=1, dtype=np.uint64)
    o = port.score_pairs(m, H, mag, ln, idx, np.full(len(idx), 7))
    assert np.abs(full["score"][idx] - o["score"]).max() <= 1e-9
    assert_flags_match(full["close"][idx], o["close"], o["score"])
    best = ctx.get_close(gm, hs, 7, hs, cand_begin=0, n_cand=n, cutoff=0.9)
    assert best[3][7] == 1            # the query itself is in the candidate range and is close to itself


def test_full_size_sweep_rows_vs_oracle(built_lib, ctx):
    """BASELINE configs[2] size (100 k x 1 kb, k=5, u8, real K1 histograms): scattered query rows of the upper-triangular sweep
    against the oracle (survivors, scores and the number of in-window pairs), and the row blocks add up to the whole sweep."""
    from meshclust2_b200 import synth
    n = 100000
    seqs, _, k, eb = synth.make_config_range("cfg3", 0, n)
    hs = ctx.count_kmers(ctx.seqs_from_text(seqs), k, eb)
    got = hs.download()
    H, mag, ln = got["hist"], got["mag"], got["len"]
    m = port.Model.from_text(weights_text("weights_cfg1_id90"))
    gm = ctx.model_from_file(weights_path("weights_cfg1_id90"))
    for q in (0, 1, 31, 4095, 50000, 50001, 99870, 99998):
        r = ctx.all_pairs(gm, hs, hs, 0.9, q_range=(q, q + 1), upper_only=True, max_out=1 << 20)
        cand = np.arange(q + 1, n)
        lo, hi = int(float(ln[q]) * 0.9), int(float(ln[q]) / 0.9)
        cand = cand[(ln[cand] >= lo) & (ln[cand] <= hi)]
        o = port.score_pairs(m, H, mag, ln, cand, np.full(len(cand), q), threads=8, want_cache=False)
        assert r["n_scored"] == len(cand), q
        want = cand[o["close"].astype(bool)]
        order = np.argsort(r["d"])
        assert np.array_equal(r["d"][order], want), q
        assert (r["q"] == q).all()
        assert np.abs(r["score"][order] - o["score"][o["close"].astype(bool)]).max(initial=0) <= 1e-9, q
    # additivity over row blocks (what the multi-GPU split relies on)
    whole = ctx.all_pairs(gm, hs, hs, 0.9, q_range=(99000, n), upper_only=True, max_out=1 << 22)
    parts = [ctx.all_pairs(gm, hs, hs, 0.9, q_range=(a, b), upper_only=True, max_out=1 << 22) for a, b in ((99000, 99333), (99333, 99334), (99334, n))]
    assert whole["n_scored"] == sum(p["n_scored"] for p in parts) and whole["n_out"] == sum(p["n_out"] for p in parts)
    key = lambda r: sorted(zip(r["q"].tolist(), r["d"].tolist(), r["score"].tolist()))
    assert key(whole) == sorted(sum((key(p) for p in parts), []))


def test_cfg4_shape_single_file_k8_u16(built_lib, ctx):
    """BASELINE configs[3] shape: records of 5 contigs x 10 kb joined by 50 N (--single-file), k=8, uint16 histograms
    (65,536 bins = 128 KiB rows): K1 through the multi-segment path, K2 through the multi-slab fast path, both vs the oracle."""
    from meshclust2_b200 import synth
    seqs = synth.make_single_file(10, 5, 10000, seed=11, n_templates=3)
    enc = built_lib.encode_batch(seqs, threads=4)
    assert (np.diff(enc["seg_off"].astype(np.int64)) == 5).all()          # the 50-N gaps split every record into 5 segments
    hs = ctx.count_kmers(ctx.upload_seqs(enc["codes"], enc["seq_off"], enc["segs"], enc["seg_off"]), 8, 2)
    got = hs.download()
    want = [port.get_point(s, 8, 2) for s in seqs]
    H = np.stack([w["hist"] for w in want])
    assert np.array_equal(got["hist"], H)
    assert np.array_equal(got["len"], np.array([w["len"] for w in want], dtype=np.uint64))
    assert np.array_equal(got["mag"], np.array([w["mag"] for w in want], dtype=np.uint64))
    rng = np.random.default_rng(5)
    mag, ln = got["mag"], got["len"]
    model = all_singles_model(FAST, H, mag, ln, rng)
    gm = ctx.model(to_desc(built_lib, model))
    ia, ib = np.repeat(np.arange(10), 10), np.tile(np.arange(10), 10)
    g = ctx.score_pairs(gm, hs, hs, ia, ib)
    o = port.score_pairs(model, H, mag, ln, ia, ib)
    assert_close_rel(g["cache"], o["cache"], 1e-8, "cfg4 cache")
    assert np.abs(g["score"] - o["score"]).max() <= 1e-8
    assert_flags_match(g["close"], o["close"], o["score"], tol=1e-8)
    # one-vs-many form (query broadcast) over 128 KiB rows
    r = ctx.score_pairs(gm, hs, hs, a_begin=0, n_pairs=10, b_begin=2, b_bc=1, want=("score",))
    assert np.abs(r["score"] - o["score"][ib == 2]).max() <= 1e-8


@pytest.mark.parametrize("k,eb", [(6, 1), (7, 2), (8, 2), (5, 2), (3, 4), (5, 1)])
def test_sweep_wide_rows_vs_oracle(built_lib, ctx, k, eb):
    """all-pairs sweep over multi-slab rows (4 KiB ... 128 KiB: query row staged in shared memory), the generic wide-type
    form and, as a control, 1 KiB rows: survivors, scores and in-window counts against the oracle; rectangular and upper."""
    rng = np.random.default_rng(50 * k + eb)
    n = 70
    hi = {1: 40, 2: 3000, 4: 100000}[eb]
    H = synth_hist(rng, n, k, eb, hi=hi)
    mag, ln = _side(H, rng, stale=True)
    ln = (ln % 200 + 900).astype(np.uint64)                       # lengths spread enough for the window to cut some pairs
    model = all_singles_model(FAST, H, mag, ln, rng)
    hs = ctx.hset_from_host(H, k, mag=mag, length=ln)
    gm = ctx.model(to_desc(built_lib, model))
    for upper, (q0, q1) in ((True, (0, n)), (False, (3, 41))):
        r = ctx.all_pairs(gm, hs, hs, 0.93, q_range=(q0, q1), upper_only=upper, max_out=n * n)
        ia, ib = [], []
        for q in range(q0, q1):
            lo, hi_ = int(float(ln[q]) * 0.93), int(float(ln[q]) / 0.93)
            for c in range(q + 1 if upper else 0, n):
                if lo <= ln[c] <= hi_:
                    ia.append(c), ib.append(q)
        o = port.score_pairs(model, H, mag, ln, ia, ib)
        assert r["n_scored"] == len(ia), (k, eb, upper)
        want = {(b, a): s for a, b, cl, s in zip(ia, ib, o["close"], o["score"]) if cl}
        got = {(q, c): s for q, c, s in zip(r["q"].tolist(), r["d"].tolist(), r["score"].tolist())}
        near = {key for key, s in zip(zip(ib, ia), o["score"]) if abs(s - 0.5) < 1e-8}
        assert set(got) - near == set(want) - near, (k, eb, upper, len(got), len(want))
        assert max([abs(got[key] - want[key]) for key in set(got) & set(want)] + [0.0]) <= 1e-8


def test_assign_rows_and_in_place_refill(built_lib, ctx, golden):
    """mc2_hset_assign_rows (batched DivergencePoint::set / clone with explicit magnitude), mc2_count_kmers_into and
    mc2_hset_update_from_device keep the side-band consistent with the bins they carry."""
    H, ln, mag = golden["hist_k5_eb1"][:12], golden["len_k5_eb1"][:12], golden["mag_k5_eb1"][:12]
    src = ctx.hset_from_host(H, 5, length=ln)
    dst = ctx.hset_from_host(np.ones((4, 1024), dtype=np.uint8), 5, length=np.ones(4, dtype=np.uint64))
    # rows 0,2 <- src 7,3 keeping dst's magnitude (set semantics); then with explicit magnitudes and lengths
    dst.assign_rows([0, 2], src, [7, 3])
    got = dst.download()
    assert np.array_equal(got["hist"][0], H[7]) and np.array_equal(got["hist"][2], H[3])
    assert got["mag"][0] == 1024 and got["len"][0] == ln[7]          # magnitude of the old all-ones row is kept
    dst.assign_rows([1, 3], src, [5, 9], mag=[111, 222], length=[1001, 1002])
    got = dst.download()
    assert list(got["mag"][[1, 3]]) == [111, 222] and list(got["len"][[1, 3]]) == [1001, 1002]
    # scoring through the assigned rows uses the carried side-band (true sums from the source rows)
    m = port.Model.from_text(weights_text("weights_appendixD_id90"))
    gm = ctx.model_from_file(weights_path("weights_appendixD_id90"))
    g = ctx.score_pairs(gm, dst, src, [1], [2], want=("score",))
    o = port.score_pairs(m, np.stack([H[5], H[2]]), np.array([111, mag[2]], dtype=np.uint64), np.array([1001, ln[2]]), [0], [1])
    assert abs(g["score"][0] - o["score"][0]) <= 1e-9
    # in-place recount equals a fresh count
    from meshclust2_b200 import synth
    seqs, _ = synth.make_set(20, 500, 3, 0.1, seed=2)
    enc = built_lib.encode_batch(seqs)
    sq = ctx.upload_seqs(enc["codes"], enc["seq_off"], enc["segs"], enc["seg_off"])
    a = ctx.count_kmers(sq, 5, 1)
    b = ctx.hset_from_host(np.ones((20, 1024), dtype=np.uint8), 5)
    ctx.count_kmers_into(sq, b)
    da, db = a.download(), b.download()
    for key in ("hist", "mag", "len", "mers1", "n_overflow"):
        assert np.array_equal(da[key], db[key]), key


def test_get_close_as_and_filter_as_equal_staged_centers(built_lib, ctx, golden):
    """mc2_get_close_as / mc2_filter_as (center = a point's row with its own magnitude and length) == staging that row into a
    scratch set with mc2_hset_assign_rows and calling mc2_get_close / mc2_filter, and == the oracle."""
    H, ln, mag = golden["hist_k5_eb1"], golden["len_k5_eb1"], golden["mag_k5_eb1"]
    cand = golden["cand"]
    hs = ctx.hset_from_host(H, 5, mag=mag, length=ln)
    sc = ctx.hset_from_host(np.ones((4, 1024), dtype=np.uint8), 5, length=np.ones(4, dtype=np.uint64))
    m = port.Model.from_text(weights_text("weights_cfg1_id90"))
    gm = ctx.model_from_file(weights_path("weights_cfg1_id90"))
    for q in (int(cand[0]), int(cand[5]), int(cand[-1])):
        for qmag, qlen in ((int(mag[q]), int(ln[q])), (int(mag[q]) + 37, int(ln[q]) - 3)):
            sc.assign_rows([1], hs, [q], mag=[qmag], length=[qlen])
            a = ctx.get_close(gm, sc, 1, hs, cand=cand, cutoff=0.9)
            b = ctx.get_close_as(gm, hs, q, qmag, qlen, hs, cand=cand, cutoff=0.9)
            assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2] and np.array_equal(a[3], b[3])
            fa = ctx.filter(gm, sc, 1, hs, cand, 0.9)
            fb = ctx.filter_as(gm, hs, q, qmag, qlen, hs, cand, 0.9)
            assert np.array_equal(fa, fb)
            # oracle: a center object = point q's bins with the overridden side-band appended as an extra row
            H2 = np.vstack([H, H[q:q + 1]])
            mag2 = np.append(mag, np.uint64(qmag)).astype(np.uint64)
            ln2 = np.append(ln, np.uint64(qlen)).astype(np.uint64)
            o = port.get_close(m, H2, mag2, ln2, len(H), cand, 0.9)
            assert b[0] == o[0] and abs(b[1] - o[1]) <= 1e-9 and b[2] == o[2] and np.array_equal(b[3], o[3])


def test_get_close_one_million_candidates_vs_oracle(built_lib, ctx):
    """BASELINE configs[4] size: Trainer<T>::get_close (src/cluster/Trainer.cpp:23-71) of one center against 10^6 candidate
    rows (1 GB of uint8 histograms, well past L2) == the oracle's get_close over the same 10^6 rows: the marks of all
    candidates, the best candidate, its distance and the is-minimum flag; contiguous range and explicit list agree."""
    rng = np.random.default_rng(21)
    n = 1000000
    base = rng.integers(1, 12, size=(4096, 1024), dtype=np.uint8)
    H = base[rng.integers(0, 4096, n)]
    rows = np.arange(n)
    for _ in range(24):                                     # 24 bins of every row nudged: related rows, few exact duplicates
        H[rows, rng.integers(0, 1024, n)] += 1
    ln = rng.integers(900, 1100, n).astype(np.uint64)
    mag = H.sum(axis=1, dtype=np.uint64)
    hs = ctx.hset_from_host(H, 5, length=ln)
    m = port.Model.from_text(weights_text("weights_cfg1_id90"))
    gm = ctx.model_from_file(weights_path("weights_cfg1_id90"))
    for q in (5, 999999):
        best, bd, ismin, marks = ctx.get_close(gm, hs, q, hs, cand_begin=0, n_cand=n, cutoff=0.9)
        ob, obd, omin, omarks = port.get_close(m, H, mag, ln, q, rows, 0.9)
        assert best == ob and ismin == omin and abs(bd - obd) <= 1e-9, q
        diff = np.flatnonzero(marks != omarks)
        if len(diff):                                       # only scores within 1e-9 of the decision boundary may differ
            o = port.score_pairs(m, H, mag, ln, diff, np.full(len(diff), q))
            assert np.abs(o["score"] - 0.5).max() <= 1e-9, q
        assert 10 < int(marks.sum()) < n // 100
        sub = np.sort(np.unique(np.concatenate([np.flatnonzero(marks), rows[::997]]))).astype(np.uint64)
        b2, bd2, min2, marks2 = ctx.get_close(gm, hs, q, hs, cand=sub, cutoff=0.9)
        assert sub[b2] == best and bd2 == bd and min2 == ismin
        assert np.array_equal(marks2, marks[sub.astype(np.int64)])
    hs.free()
