"""Shared test helpers: synthetic inputs, hand-made models, comparison utilities."""
import numpy as np

from oracle import port

FAST = [port.FEAT[n] for n in port.FAST]
SLOW = [port.FEAT[n] for n in port.SLOW]
REL_TOL = 1e-9   # floating-point singles / scores: CUDA vs oracle (north_star asks 1e-6; we hold 1e-9)
SCORE_ATOL = 1e-9


def synth_hist(rng, n, k, eb, hi=None, related=True):
    """n histograms of 4^k bins (dtype by eb), values >= 1; related=True makes near-duplicates so some pairs are close."""
    N = 4 ** k
    dt = port.DTYPES[eb]
    top = int(np.iinfo(dt).max) if hi is None else hi
    base = rng.integers(1, min(top, 12) + 1, size=(max(1, n // 8), N))
    H = base[rng.integers(0, len(base), n)]
    if related:
        noise = rng.integers(-1, 2, size=H.shape) * (rng.random(H.shape) < 0.15)
        H = H + noise
    H = np.clip(H, 1, top).astype(dt)
    return H


def all_singles_model(flags, H, mag, ln, rng, n_combos=6, bias=0.0):
    """A model over the given single flags with min/max measured on random pairs (like Feature::normalize) and random
    combos/weights; returns oracle.port.Model."""
    n = H.shape[0]
    ia, ib = rng.integers(0, n, 64), rng.integers(0, n, 64)
    singles = []
    for f in flags:
        vals = []
        for a, b in zip(ia, ib):
            vals.append(port.raw_single(f, H[a], H[b], int(mag[a]), int(mag[b]), int(ln[a]), int(ln[b])))
        lo, hi = float(np.nanmin(vals)), float(np.nanmax(vals))
        if not np.isfinite(lo) or not np.isfinite(hi) or hi - lo < 1e-9:
            lo, hi = lo if np.isfinite(lo) else 0.0, (lo if np.isfinite(lo) else 0.0) + 1.0
        singles.append((f, lo, hi))
    combos = []
    for c in range(n_combos):
        kind = int(rng.integers(0, 4))
        i, j = sorted(rng.choice(len(flags), size=2, replace=False).tolist())
        fl = flags[i] | flags[j]
        if kind in (0, 3) and rng.random() < 0.3:
            fl = flags[i]
        combos.append((kind, fl))
    # every single must appear in some combo (the weights-file lookup order is derived from the combos)
    used = 0
    for _, fl in combos:
        used |= fl
    for f in flags:
        if not used & f:
            combos.append((0, f))
    # order singles as add_feature would see them
    look = []
    for _, fl in combos:
        for b in range(64):
            if fl >> b & 1 and (1 << b) not in look:
                look.append(1 << b)
    norm = {f: (lo, hi) for f, lo, hi in singles}
    singles = [(f,) + norm[f] for f in look]
    weights = (rng.standard_normal(len(combos) + 1) * 2).tolist()
    return port.Model(singles, combos, weights, bias, k=5, ident=0.9, datatype="uint8_t")


def to_desc(capi, model, regression=0):
    """oracle.port.Model -> capi.ModelDesc (the product's own struct)"""
    singles = model.singles
    combos = [(kind, idx) for (kind, _), idx in zip(model.combos, model.combo_indices())]
    return capi.make_desc(singles, combos, model.weights, model.bias, regression)


def assert_close_rel(a, b, rel=REL_TOL, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    both_nan = np.isnan(a) & np.isnan(b)
    denom = np.maximum(np.abs(b), 1e-300)
    err = np.where(both_nan, 0.0, np.abs(a - b) / denom)
    err = np.where((a == b), 0.0, err)
    worst = np.nanmax(err) if err.size else 0.0
    assert not np.isnan(err).any() and worst <= rel, "%s: max rel err %.3e" % (what, worst)


def assert_flags_match(close_a, close_b, score_b, bias=0.0, tol=1e-9, what=""):
    """close flags must be identical except where the oracle's score is within tol of the 0.5 decision boundary"""
    diff = np.nonzero(np.asarray(close_a) != np.asarray(close_b))[0]
    for j in diff:
        assert abs(score_b[j] - 0.5) <= tol, "%s: close flag differs at %d with score %.17g" % (what, j, score_b[j])
