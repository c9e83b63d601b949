"""GPU: the input contract on the device (mc2_seqs_from_text: Chromosome::help + ChromosomeOneDigit::encode) against the
oracle's encode (pinned to the reference in tests/test_oracle_vs_ref.py): segment lists bit-exact, and the histograms counted
from the device-built sequences equal the oracle's."""
import numpy as np
import pytest

from oracle import port
from meshclust2_b200 import synth

pytestmark = pytest.mark.gpu


def _oracle_segments(seqs):
    segs, off = [], [0]
    for s in seqs:
        _, sg, _ = port.encode(s)
        segs += [tuple(x) for x in np.asarray(sg).reshape(-1, 2).tolist()]
        off.append(len(segs))
    return np.array(segs, dtype=np.int32).reshape(-1, 2), np.array(off, dtype=np.uint64)


def _adversarial(rng):
    def rnd(n, alphabet=b"ACGT"):
        return bytes(rng.choice(np.frombuffer(alphabet, dtype=np.uint8), n))
    cases = [
        rnd(1000), rnd(31), rnd(32), rnd(33), rnd(64), rnd(20), rnd(21), rnd(19), rnd(1), b"",
        b"N" * 50, b"N" * 50 + b"A", b"N" * 50 + b"AC",                      # a run opening on the last base is dropped (Q7)
        rnd(40) + b"N" * 9 + rnd(40), rnd(40) + b"N" * 10 + rnd(40),         # gap 9: next.s - cur.e = 10 -> not bridged
        rnd(40) + b"N" * 8 + rnd(40),                                        # gap 8: bridged, N -> C inside the segment
        rnd(10) + b"N" * 3 + rnd(5) + b"N" * 30 + rnd(25),                   # merged 18 < 20 dropped, 25 kept
        rnd(19) + b"N" * 12 + rnd(19), rnd(5) + b"N" + rnd(5),               # all dropped / short sequence keeps everything
        rnd(300, b"acgtn"), rnd(300, b"ACGTRYMKSWHBVDX"), rnd(200, b"ACGTNNNN"),
        b"N" + rnd(63) + b"N", rnd(31) + b"N" + rnd(32), rnd(32) + b"N" * 32 + rnd(32),
        b"ACGT" * 8 + b"N" * 20 + b"ACGT" * 8 + b"N",
        rnd(1000001) + b"N" * 10 + rnd(2500000) + b"N" * 3 + rnd(30),          # 1 Mbp splitting, last piece takes the rest
    ]
    return cases


def test_segments_equal_the_oracle(built_lib, ctx, golden_seqs):
    rng = np.random.default_rng(17)
    seqs = list(golden_seqs) + _adversarial(rng)
    # golden_seqs hold deliberate invalid letters? keep only those the oracle encodes
    ok = []
    for s in seqs:
        try:
            port.encode(s)
            ok.append(s)
        except ValueError:
            pass
    assert len(ok) >= len(seqs) - 3
    sq = ctx.seqs_from_text(ok)
    segs, off, ln = sq.segments()
    wsegs, woff = _oracle_segments(ok)
    assert np.array_equal(off, woff)
    assert np.array_equal(segs, wsegs)
    assert ln.tolist() == [len(s) for s in ok]


@pytest.mark.parametrize("k,eb", [(5, 1), (3, 2), (8, 2)])
def test_histograms_from_device_built_sequences(built_lib, ctx, golden_seqs, k, eb):
    rng = np.random.default_rng(k)
    seqs = [s for s in list(golden_seqs) + _adversarial(rng)[:-1] if len(s) > 0]
    good = []
    for s in seqs:
        try:
            port.encode(s)
            good.append(s)
        except ValueError:
            pass
    hs = ctx.count_kmers(ctx.seqs_from_text(good), k, eb)
    got = hs.download()
    for i, s in enumerate(good):
        w = port.get_point(s, k, eb)
        assert np.array_equal(got["hist"][i], w["hist"]), (k, eb, i)
        assert np.array_equal(got["mers1"][i], w["mers1"]) and got["len"][i] == w["len"] and got["mag"][i] == w["mag"], (k, eb, i)
    # same as the host path
    enc = built_lib.encode_batch(good, threads=2)
    hs2 = ctx.count_kmers(ctx.upload_seqs(enc["codes"], enc["seq_off"], enc["segs"], enc["seg_off"]), k, eb)
    assert np.array_equal(hs2.download()["hist"], got["hist"])


def test_invalid_letter_and_empty_batch(built_lib, ctx):
    with pytest.raises(built_lib.Mc2Error):
        ctx.seqs_from_text([b"ACGT" * 10 + b"!" + b"ACGT" * 10])
    with pytest.raises(built_lib.Mc2Error):
        ctx.seqs_from_text([b"ACGT" * 10 + b"N" * 30 + b"ACJT" * 2])       # invalid letter outside every segment, but segments exist
    sq = ctx.seqs_from_text([b"N" * 30 + b"!A"])                            # no segment at all: nothing is encoded, no error
    assert sq.segments()[0].shape[0] == 0
    assert len(ctx.seqs_from_text([])) == 0
    # 100k x 1 kb: the device path agrees with the host path on every segment
    seqs, _, k, eb = synth.make_config_range("cfg3", 0, 20000)
    segs, off, _ = ctx.seqs_from_text(seqs).segments()
    enc = built_lib.encode_batch(seqs, threads=8)
    assert np.array_equal(off, enc["seg_off"]) and np.array_equal(segs, np.asarray(enc["segs"]).reshape(-1, 2))
