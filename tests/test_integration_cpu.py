"""CPU, end to end: the HOST logic of the relinked meshclust2 (integration/*.cpp + the build-time hunks: FASTA record
splitting, find_k / width detection / get_points offers, point objects, batched update and merge passes, one-call-per-center
paths) with the device library replaced by a stand-in that serves the same C ABI from the CPU oracle
(tests/cpp/stub_device_oracle.cpp, `make -C oracle integrated_stub`).  The binary must print the same decisions and produce the
same weights.txt and clusters as the unmodified reference binary in every form.  This does NOT test the CUDA kernels -- that
is tests/test_integrated_cluster.py on the GPU, same assertions, real library.
Skipped where the reference sources or oracle/_ref are absent."""
import os
import subprocess

import pytest

from conftest import ROOT
from meshclust2_b200 import synth
from test_integrated_cluster import parse_clstr, run

REF_ROOT = os.environ.get("MC2_REFERENCE_ROOT", "/root/reference")
REF = os.path.join(ROOT, "oracle", "_ref", "meshclust2")
STUB = os.path.join(ROOT, "oracle", "_ref", "meshclust2_stub")

pytestmark = pytest.mark.skipif(not (os.path.isdir(os.path.join(REF_ROOT, "src", "cluster")) and os.path.exists(REF)),
                                reason="reference sources / oracle/_ref not present")


@pytest.fixture(scope="module")
def stub_binary():
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "oracle"), "integrated_stub", "REF=" + REF_ROOT])
    return STUB


@pytest.mark.timeout(600)
def test_host_logic_reproduces_the_reference(stub_binary, tmp_path):
    seqs, tids = synth.make_set(500, 1000, 80, 0.08, seed=11)
    seqs = list(seqs)
    a = bytearray(seqs[3]); a[200:240] = b"N" * 40; a[500:504] = b"nnnn"; seqs[3] = bytes(a)   # split + bridged gap
    seqs[4] = seqs[4].lower()
    fasta = str(tmp_path / "in.fa")
    open(fasta, "w").write(synth.to_fasta(seqs, tids))
    w_ref, d_ref = run(REF, fasta, str(tmp_path / "ref"), str(tmp_path / "ref.clstr"))
    c_ref = parse_clstr(str(tmp_path / "ref.clstr"))
    assert len(d_ref) == 4 and sum(len(c) for c in c_ref) == 500 and 30 < len(c_ref) < 500
    for tag, env in (("batched", {}), ("ingest", {"MC2_K1_MIN_BASES": "0"}),
                     ("k1", {"MC2_K1_MIN_BASES": "0", "MC2_NO_DEVICE_READER": "1"}), ("percall", {"MC2_NO_BATCH": "1"})):
        w, d = run(stub_binary, fasta, str(tmp_path / tag), str(tmp_path / (tag + ".clstr")), env)
        assert d == d_ref, (tag, d_ref, d)
        assert w == w_ref, tag
        c = parse_clstr(str(tmp_path / (tag + ".clstr")))
        assert c == c_ref, "%s: %d vs %d clusters, %d in common" % (tag, len(c_ref), len(c), len(c_ref & c))


@pytest.mark.timeout(600)
def test_no_train_list_host_logic(stub_binary, tmp_path):
    """--no-train-list (src/cluster/CRunner.cpp:576-592): points appended and every id re-assigned after the Trainer was
    built; the integration's id-keyed device mirror must notice and refill itself (integration/Trainer_b200.cpp rows_of)"""
    seqs, tids = synth.make_set(400, 1000, 60, 0.08, seed=17)
    train, extra = str(tmp_path / "train.fa"), str(tmp_path / "extra.fa")
    open(train, "w").write(synth.to_fasta(seqs[:250], tids[:250]))
    open(extra, "w").write("".join(">x%d template_%d\n%s\n" % (i, tids[250 + i], s.decode()) for i, s in enumerate(seqs[250:])))
    lst = str(tmp_path / "list.txt")
    open(lst, "w").write(extra + "\n")
    args = ("--no-train-list", lst)
    w_ref, _ = run(REF, train, str(tmp_path / "ref"), str(tmp_path / "ref.clstr"), extra=args)
    c_ref = parse_clstr(str(tmp_path / "ref.clstr"))
    assert sum(len(c) for c in c_ref) == 400
    for tag, env in (("batched", {}), ("percall", {"MC2_NO_BATCH": "1"}), ("noprewarm", {"MC2_NO_PREWARM": "1"})):
        w, _ = run(stub_binary, train, str(tmp_path / tag), str(tmp_path / (tag + ".clstr")), env, extra=args)
        assert w == w_ref, tag
        c = parse_clstr(str(tmp_path / (tag + ".clstr")))
        assert c == c_ref, "%s: %d vs %d clusters, %d in common" % (tag, len(c_ref), len(c), len(c_ref & c))
