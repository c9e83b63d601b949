"""CPU: the C restatement (oracle/mc2_oracle.c) against the committed golden vectors produced by the reference itself
(tests/golden/make_golden.py) — this is what pins the oracle."""
import numpy as np
import pytest

from conftest import weights_text
from oracle import port


def test_encode_and_segments(golden, golden_seqs):
    off, soff = golden["text_off"], golden["seg_off"]
    for i, s in enumerate(golden_seqs):
        codes, segs, eff = port.encode(s)
        assert np.array_equal(codes, golden["codes"][off[i]:off[i + 1]]), i
        assert np.array_equal(segs.reshape(-1, 2), golden["segs"][soff[i]:soff[i + 1]]), i
        assert eff == golden["eff"][i]


@pytest.mark.parametrize("k,eb", [(5, 1), (3, 2), (2, 4), (4, 8), (6, 1), (3, 1)])
def test_histograms_bit_exact(golden, golden_seqs, k, eb):
    tag = "k%d_eb%d" % (k, eb)
    for i, s in enumerate(golden_seqs):
        p = port.get_point(s, k, eb)
        assert np.array_equal(p["hist"], golden["hist_" + tag][i]), (tag, i)
        assert np.array_equal(p["mers1"], golden["mers1_" + tag][i])
        assert p["mag"] == golden["mag_" + tag][i] and p["len"] == golden["len_" + tag][i]
        assert abs(p["stddev"] - golden["stddev_" + tag][i]) <= 1e-12 * max(1.0, golden["stddev_" + tag][i])


def test_saturation_case_present(golden):
    # the poly-A sequence saturates uint8 at k=3: the fixture must exercise the overflow branch
    assert golden["hist_k3_eb1"].max() == 255


@pytest.mark.parametrize("k,eb", [(5, 1), (3, 2), (2, 4), (4, 8)])
def test_raw_singles(golden, k, eb):
    tag = "k%d_eb%d" % (k, eb)
    H, ln = golden["hist_" + tag], golden["len_" + tag]
    names = [str(x) for x in golden["single_names"]]
    ia, ib = golden["pair_ia"], golden["pair_ib"]
    for j in range(len(ia)):
        for c, nm in enumerate(names):
            want = golden["raw_" + tag][j, c]
            la, lb = int(ln[ia[j]]), int(ln[ib[j]])
            if nm == "length_difference" and (la == 0 or lb == 0):
                with pytest.raises(ValueError):
                    port.raw_single(port.FEAT[nm], H[ia[j]], H[ib[j]], None, None, la, lb)
                continue
            got = port.raw_single(port.FEAT[nm], H[ia[j]], H[ib[j]], None, None, la, lb)
            if np.isnan(want):
                assert np.isnan(got)
            elif nm in ("manhattan", "emd", "length_difference"):
                assert got == want, (nm, j)           # integer-valued singles: bit exact
            else:
                assert abs(got - want) <= 1e-12 * max(abs(want), 1e-300), (nm, j, got, want)


def test_raw_singles_stale_mag(golden):
    H, ln, mag = golden["hist_k5_eb1"], golden["len_k5_eb1"], golden["stale_mag_k5_eb1"]
    names = [str(x) for x in golden["single_names"]]
    ia, ib = golden["pair_ia"], golden["pair_ib"]
    for j in range(len(ia)):
        for c, nm in enumerate(names):
            want = golden["raw_stale_k5_eb1"][j, c]
            la, lb = int(ln[ia[j]]), int(ln[ib[j]])
            if np.isnan(want):
                continue
            got = port.raw_single(port.FEAT[nm], H[ia[j]], H[ib[j]], int(mag[ia[j]]), int(mag[ib[j]]), la, lb)
            assert abs(got - want) <= 1e-12 * max(abs(want), 1e-300), (nm, j, got, want)


@pytest.mark.parametrize("wname", ["weights_cfg1_id90", "weights_appendixD_id90"])
def test_classifier_and_callers(golden, wname):
    m = port.Model.from_text(weights_text(wname))
    H, ln, mag = golden["hist_k5_eb1"], golden["len_k5_eb1"], golden["mag_k5_eb1"]
    ja, jb = golden["score_ia"], golden["score_ib"]
    r = port.score_pairs(m, H, mag, ln, ja, jb)
    assert np.array_equal(r["close"], golden[wname + "_close"])
    assert np.abs(r["score"] - golden[wname + "_score"]).max() <= 1e-12
    assert np.abs(r["dist"] - golden[wname + "_dist"]).max() <= 1e-12
    assert np.abs(r["cache"] - golden[wname + "_cache"]).max() <= 1e-12
    r2 = port.score_pairs(m, H, golden["stale_mag_k5_eb1"], ln, ja, jb)
    assert np.array_equal(r2["close"], golden[wname + "_stale_close"])
    assert np.abs(r2["score"] - golden[wname + "_stale_score"]).max() <= 1e-12
    cand = golden["cand"]
    for t, q in enumerate(golden[wname + "_gc_q"]):
        best, bd, ismin, marks = port.get_close(m, H, mag, ln, int(q), cand, 0.9)
        assert best == golden[wname + "_gc_best"][t]
        assert abs(bd - golden[wname + "_gc_dist"][t]) <= 1e-12
        assert ismin == bool(golden[wname + "_gc_ismin"][t])
        assert np.array_equal(marks, golden[wname + "_gc_marks"][t])
        keep = port.filter_members(m, H, mag, ln, int(q), cand, 0.9)
        assert np.array_equal(keep, golden[wname + "_filter_keep"][t])
        rows = cand[(np.arange(8) + int(q)) % len(cand)]
        assert port.merge(m, H, mag, ln, rows, 0, 1, 7, 0.9) == golden[wname + "_merge"][t]


def test_distance(golden):
    H = golden["hist_k5_eb1"]
    ia, ib = golden["pair_ia"], golden["pair_ib"]
    got = np.array([port.distance(H[a], H[b]) for a, b in zip(ia, ib)], dtype=np.uint64)
    assert np.array_equal(got, golden["distance_k5_eb1"])
    C = golden["distance_d_centers"]
    gd = np.array([port.distance_d(H[ib[j]], C[j]) for j in range(40)])
    assert np.abs(gd - golden["distance_d_k5_eb1"]).max() <= 1e-9


def test_weights_roundtrip():
    for wname in ("weights_cfg1_id90", "weights_appendixD_id90"):
        txt = weights_text(wname)
        m = port.Model.from_text(txt)
        assert port.Model.from_text(m.to_text()).to_text() == m.to_text()
        assert [x.split() for x in m.to_text().strip().splitlines() if x.strip()] == \
               [x.split() for x in txt.strip().splitlines() if x.strip()]
