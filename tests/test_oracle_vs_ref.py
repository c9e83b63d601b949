"""CPU, build container only: the C restatement against the reference itself (oracle/_ref/libmc2ref.so compiled from
/root/reference by oracle/Makefile) on fresh random inputs.  Skipped where the reference library is not present."""
import numpy as np
import pytest

from conftest import weights_text
from oracle import port, ref
from meshclust2_b200 import synth

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")


def test_get_point_random_and_adversarial():
    rng = np.random.default_rng(5)
    seqs, _ = synth.make_set(24, 400, 4, 0.1, seed=17)

    def adv(s):
        a = bytearray(s)
        for _ in range(int(rng.integers(0, 4))):
            p, L = int(rng.integers(0, len(a))), int(rng.integers(1, 60))
            a[p:p + L] = b"N" * len(a[p:p + L])
        for _ in range(4):
            a[int(rng.integers(0, len(a)))] = int(rng.choice(list(b"RYMKSWHBVDacgtn")))
        return bytes(a)
    tests = [adv(s) for s in seqs] + [b"A", b"AN", b"NA", b"ACGTN", b"NNNN", b"ACGTACGTACGTACGTACGTA"]
    for s in tests:
        assert all(np.array_equal(x, y) if isinstance(x, np.ndarray) else x == y
                   for x, y in zip(ref.encode(s), port.encode(s)))
        for k, eb in ((1, 1), (3, 1), (5, 2), (4, 4), (2, 8)):
            r, p = ref.get_point(s, k, eb), port.get_point(s, k, eb)
            assert np.array_equal(r["hist"], p["hist"]) and np.array_equal(r["mers1"], p["mers1"])
            assert r["mag"] == p["mag"] and r["len"] == p["len"]
            assert abs(r["stddev"] - p["stddev"]) <= 1e-12 * max(1.0, abs(r["stddev"]))


def test_invalid_letter_throws_in_both():
    s = b"ACGTACGTACGTACGTACGTACGTJACGTACGT"
    with pytest.raises(ValueError):
        ref.encode(s)
    with pytest.raises(ValueError):
        port.encode(s)


@pytest.mark.parametrize("eb", [1, 2, 4, 8])
def test_raw_singles_all_widths(eb):
    rng = np.random.default_rng(eb)
    dt = port.DTYPES[eb]
    worst = 0.0
    for trial in range(12):
        k = int(rng.integers(1, 6))
        N = 4 ** k
        hi = [6, min(int(np.iinfo(dt).max), 70000), 300, 3][trial % 4]
        hi = min(hi, int(np.iinfo(dt).max))
        p = rng.integers(1, hi + 1, size=N).astype(dt)
        q = rng.integers(1, hi + 1, size=N).astype(dt)
        for name in port.SLOW:
            lp, lq = int(rng.integers(1, 2000)), int(rng.integers(1, 2000))
            r = ref.raw_single(port.FEAT[name], p, q, 0, 0, lp, lq, k=k)
            o = port.raw_single(port.FEAT[name], p, q, None, None, lp, lq)
            if r == o or (np.isnan(r) and np.isnan(o)):
                continue
            worst = max(worst, abs(r - o) / max(abs(r), 1e-300))
    assert worst <= 1e-12   # reference build uses fused multiply-adds; the restatement does not


@pytest.mark.parametrize("wname", ["weights_cfg1_id90", "weights_appendixD_id90"])
def test_trainer_callers(wname):
    seqs, _ = synth.make_set(120, 1000, 10, 0.12, seed=23)
    pts = [port.get_point(s, 5, 1) for s in seqs]
    H = np.stack([p["hist"] for p in pts])
    mag = np.array([p["mag"] for p in pts], dtype=np.uint64)
    ln = np.array([p["len"] for p in pts], dtype=np.uint64)
    txt = weights_text(wname)
    m, rm = port.Model.from_text(txt), ref.RefModel(txt, 1, 0.9)
    rng = np.random.default_rng(3)
    ia, ib = rng.integers(0, 120, 1500), rng.integers(0, 120, 1500)
    o = port.score_pairs(m, H, mag, ln, ia, ib)
    r = rm.score_pairs(H, None, ln, ia, ib, mode=0, n_singles=len(m.singles))
    assert np.array_equal(o["close"], r["close"]) and o["close"].any() and not o["close"].all()
    assert np.abs(o["score"] - r["score"]).max() <= 1e-12
    assert np.array_equal(o["close"], rm.score_pairs(H, None, ln, ia, ib, mode=1)["close"])   # Predictor::close
    for q in range(0, 120, 17):
        cand = np.array([c for c in range(120) if c != q])
        ob, rb = port.get_close(m, H, mag, ln, q, cand, 0.9), rm.get_close(H, None, ln, q, cand)
        assert ob[0] == rb[0] and ob[2] == rb[2] and np.array_equal(ob[3], rb[3]) and abs(ob[1] - rb[1]) <= 1e-12
        assert np.array_equal(port.filter_members(m, H, mag, ln, q, cand, 0.9), rm.filter_members(H, None, ln, q, cand))
        rows = np.arange(q, min(q + 8, 120))
        assert port.merge(m, H, mag, ln, rows, 0, 1, len(rows) - 1, 0.9) == rm.merge(H, None, ln, rows, 0, 1, len(rows) - 1)


@pytest.mark.parametrize("eb", [1, 2, 4, 8])
def test_k3_mean_closest(eb):
    """get_mean / mean_shift_update mean + Trainer::closest: same arg-min and bit-identical mean; distances within an ulp
    (the reference build fuses 1 - frac*frac)"""
    rng = np.random.default_rng(40 + eb)
    for trial in range(15):
        N = 4 ** int(rng.integers(1, 6))
        hi = {1: 255, 2: 3000, 4: 100000, 8: 100000}[eb]
        H = rng.integers(1, hi + 1, size=(50, N)).astype(port.DTYPES[eb])
        mem = rng.integers(0, 50, int(rng.integers(1, 40)))
        a, b = port.mean_closest(H, mem), ref.mean_closest(H, mem)
        assert a[0] == b[0] and np.array_equal(a[2], b[2])
        assert np.abs(a[3] - b[3]).max() <= 1e-12 * max(1.0, np.abs(b[3]).max())


def test_width_detection_largest_count():
    """Runner::run's "Largest count" (CRunner.cpp:57-93) and the width rule (CRunner.cpp:108-126)"""
    from meshclust2_b200 import synth
    seqs, _ = synth.make_range(40, 1000, 5, 0.1, seed=3)
    sets = {"plain": seqs, "homopolymer": seqs + [b"A" * 700], "segments": seqs + [b"ACGT" * 100 + b"N" * 30 + b"ACGT" * 90],
            "u16": seqs + [b"AC" * 40000], "one": [b"ACGTTGCAAGGCTTAACCGGTTAAC"]}
    for name, ss in sets.items():
        for k in (2, 5, 8):
            want = ref.largest_count(ss, k)
            got = 0
            for s in ss:
                c, sg, _ = port.encode(s)
                got = max(got, port.largest_count(c, sg, k))
            assert got == want, (name, k)
    assert [port.width_for(v) for v in (0, 255, 256, 65535, 65536, 2 ** 32 - 1, 2 ** 32)] == [1, 1, 2, 2, 4, 4, 8]
    # a segment shorter than k: not restated (the reference reads past the segment)
    c, sg, _ = port.encode(b"ACGTAC")
    with pytest.raises(ValueError):
        port.largest_count(c, sg, 8)
