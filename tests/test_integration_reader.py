"""CPU: the host-side pieces of the relinked binary's device reader (integration/fasta_records.h: FASTA record splitting,
the encoded sequence string, the doubled lengths of Runner::find_k) against the reference's own reader
(ChromListMaker / Chromosome / ChromosomeOneDigitDna, linked from oracle/_ref/libmc2ref.so).  The device half of the reader
(segments, histograms) is covered by tests/test_gpu_ingest.py and tests/test_integrated_cluster.py.
Skipped where the reference sources or oracle/_ref are absent (the GPU box has no /root/reference)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

REF = os.environ.get("MC2_REFERENCE_ROOT", "/root/reference")
LIBREF = os.path.join(ROOT, "oracle", "_ref", "libmc2ref.so")
pytestmark = pytest.mark.skipif(not (os.path.isdir(os.path.join(REF, "src", "nonltr")) and os.path.exists(LIBREF)),
                                reason="reference sources / oracle/_ref not present")


def _compile(tmp_path_factory, name):
    exe = str(tmp_path_factory.mktemp(name) / name)
    inc = ["-I" + os.path.join(REF, "src", d) for d in ("", "clutil", "predict", "nonltr", "utility", "exception", "cluster")]
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-fopenmp", "-pthread", "-w", "-include", "cstdint", "-include", "limits"] + inc +
                          ["-I" + os.path.join(ROOT, "integration"), "-o", exe, os.path.join(ROOT, "tests", "cpp", name + ".cpp"),
                           "-L" + os.path.dirname(LIBREF), "-lmc2ref", "-Wl,-rpath," + os.path.dirname(LIBREF)])
    return exe


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    return _compile(tmp_path_factory, "test_fasta_records")


@pytest.fixture(scope="module")
def points_harness(tmp_path_factory):
    return _compile(tmp_path_factory, "test_build_points")


def _run(exe, path):
    r = subprocess.run([exe, path], capture_output=True, text=True, timeout=120)
    return r.returncode, r.stdout.strip()


def _dna(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(list(alphabet), size=n).tolist())


def test_reader_pieces_against_the_reference(harness, tmp_path):
    rng = np.random.default_rng(3)
    recs = []
    recs.append(("plain 70-column", _dna(rng, 1000)))
    recs.append(("lower case", _dna(rng, 300).lower()))
    s = list(_dna(rng, 900))
    s[100:130] = "N" * 30; s[300:305] = "N" * 5; s[500:512] = "n" * 12; s[520:535] = "N" * 15     # long, bridged and short-gap runs
    recs.append(("N runs", "".join(s)))
    recs.append(("N at both ends", "NNNNNNNNNNNNNNNNNNNNNNNN" + _dna(rng, 200) + "NNNNNNNNNNNNNNNNNNNNNNNNNNNNNNN"))
    recs.append(("lone base after N", "ACGTACGTACGTACGTACGTACGTACGT" + "N" * 25 + "A"))
    recs.append(("short island", _dna(rng, 100) + "N" * 30 + "ACGTACGTAC" + "N" * 30 + _dna(rng, 100)))
    recs.append(("IUPAC", _dna(rng, 400, "ACGTRYMKSWHBVDX")))
    recs.append(("all N", "N" * 60))
    recs.append(("tiny", "ACGTAC"))
    text = ""
    for j, (name, seq) in enumerate(recs):
        text += ">rec%d %s\n" % (j, name)
        w = 70 if j % 2 == 0 else 61
        text += "\n".join(seq[i:i + w] for i in range(0, len(seq), w)) + "\n"
        if j == 3:
            text += "\n \tskipped line starting with a blank\n\tanother skipped line\n"
    for tag, body in (("lf", text), ("crlf", text.replace("\n", "\r\n")), ("cr", text.replace("\n", "\r")),
                      ("no_final_newline", text.rstrip("\n"))):
        p = tmp_path / (tag + ".fa")
        p.write_bytes(body.encode())
        rc, out = _run(harness, str(p))
        assert rc == 0 and out == "OK %d records" % len(recs), (tag, out)


def test_invalid_letters_are_rejected_like_the_reference(harness, tmp_path):
    p = tmp_path / "bad.fa"
    p.write_text(">ok\nACGTACGTACGTACGTACGTACGTACGTACGT\n>bad\nACGTACGTACGTAC*TACGTACGTACGTACGTACGT\n")
    rc, out = _run(harness, str(p))
    assert rc == 0 and out == "OK both reject", out
    p = tmp_path / "bad_outside.fa"                 # an invalid letter outside every segment is rejected too (the second loop
    p.write_text(">bad\n" + "ACGT" * 20 + "N" * 30 + "AC?GT" + "N" * 30 + "ACGT" * 20 + "\n")   # of ChromosomeOneDigit::encode)
    rc, out = _run(harness, str(p))
    assert rc == 0 and out == "OK both reject", out


def test_corner_shapes_are_declined(harness, tmp_path):
    for name, body in (("no_header", "ACGTACGT\n>late\nACGT\n"), ("empty_record", ">a\n>b\nACGTACGT\n"), ("nothing", "\n\n")):
        p = tmp_path / (name + ".fa")
        p.write_text(body)
        rc, out = _run(harness, str(p))
        assert rc == 0 and out == "DECLINED", (name, out)


def test_host_points_equal_loader_get_point(points_harness, tmp_path):
    """build_points (the host objects of the device reader, std::thread workers) == Loader<T>::get_point field by field, for
    uint8 and uint16 histograms, 1 and 7 worker threads, more records than one work block"""
    rng = np.random.default_rng(11)
    text = ""
    for j in range(700):
        s = list(_dna(rng, int(rng.integers(40, 400))))
        if j % 9 == 0:
            a = int(rng.integers(0, len(s) - 30)); s[a:a + 25] = "N" * 25
        if j % 13 == 0:
            s[:5] = "nnnnn"
        if j % 17 == 0:
            s = list("A" * 700)                          # saturates uint8 bins
        text += ">read_%d some description\n" % j + "".join(s) + "\n"
    p = tmp_path / "many.fa"
    p.write_text(text)
    for k, threads in ((3, 1), (4, 7)):
        r = subprocess.run([points_harness, str(p), str(k), str(threads)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and r.stdout.split() == ["OK", "700", "points", "OK", "700", "points"], r.stdout + r.stderr
