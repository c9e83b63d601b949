"""End to end: the reference's own fastcar with ONE call added to work() (src/fastcar/FC_Runner.cpp:427-470 ->
integration/patch_fc_runner.py + integration/FastcarWork_b200.cpp: each query-chunk x database-chunk block goes to the device
as one mc2_all_pairs + one mc2_score_pairs) must write the same output lines as the unmodified reference binary.

Models come from weights files (--recover): the reference's in-process training leaves Feature's memo table switched on
(Feature::get_func, src/predict/Feature.cpp:460-465) and fastcar re-uses point ids between the training set, the queries and
every database chunk, so a freshly trained fastcar answers some pairs from stale memo entries -- a reference defect the device
path does not imitate (documented in INTEGRATION.md).  The default "rc" training also crashes in this build of the reference.

GPU form: oracle/_ref/fastcar_b200 (real library).  CPU form: oracle/_ref/fastcar_stub (the C ABI served by the oracle) checks
the host logic -- block handling, the reference's bin_search quirk, output order -- where no GPU is present."""
import os
import subprocess

import pytest

from conftest import ROOT
from meshclust2_b200 import synth

REF_ROOT = os.environ.get("MC2_REFERENCE_ROOT", "/root/reference")
REF = os.path.join(ROOT, "oracle", "_ref", "fastcar")
OURS = os.path.join(ROOT, "oracle", "_ref", "fastcar_b200")
STUB = os.path.join(ROOT, "oracle", "_ref", "fastcar_stub")

REGRESSION_SECTION = """n_combos: 3
0.35
0 32 0.5
0 8 0.1
3 262144 0.05

n_singles: 3
32 0.793168111824964 0.99479796935211
8 7.14142842854285 47.549973711875
262144 1052 282849
"""


def _inputs(tmp_path):
    seqs, tids = synth.make_set(600, 1000, 60, 0.08, seed=23)
    seqs = list(seqs)
    seqs[7] = seqs[7][:700]            # lengths outside most windows: exercises work()'s start index (bin_search quirk)
    seqs[9] = seqs[9][:640]
    seqs[460] = seqs[460][:800]
    db, q = str(tmp_path / "db.fa"), str(tmp_path / "q.fa")
    open(db, "w").write(synth.to_fasta(seqs[:450], tids[:450]))
    open(q, "w").write(synth.to_fasta(seqs[450:], tids[450:]))
    w1 = os.path.join(ROOT, "tests", "golden", "weights_cfg1_id90.txt")
    w3 = str(tmp_path / "w3.txt")
    open(w3, "w").write(open(w1).read().replace("mode: 1", "mode: 3").rstrip("\n") + "\n\n" + REGRESSION_SECTION)
    return db, q, w1, w3


def _run(binary, db, q, wd, args, env=None):
    os.makedirs(wd, exist_ok=True)
    r = subprocess.run([binary, db, "--query", q, "--id", "0.9", "--threads", "1", "--output", os.path.join(wd, "out")] + args,
                       cwd=wd, capture_output=True, text=True, timeout=900, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    n_pos = [l for l in r.stdout.splitlines() if "predicted positive" in l]
    return open(os.path.join(wd, "out0")).read(), n_pos


def _compare(binary, tmp_path):
    db, q, w1, w3 = _inputs(tmp_path)
    for tag, args in (("class", ["--recover", w1]), ("class_regr", ["--recover", w3]), ("chunks", ["--recover", w3, "--chunk", "64"]),
                      ("full_header", ["--recover", w3, "--no-format"])):
        want, pos_ref = _run(REF, db, q, str(tmp_path / ("ref_" + tag)), args)
        got, pos = _run(binary, db, q, str(tmp_path / ("our_" + tag)), args)
        assert len(want.splitlines()) > 100
        assert got == want, "%s: %d vs %d lines" % (tag, len(got.splitlines()), len(want.splitlines()))
        assert pos == pos_ref


@pytest.mark.gpu
@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(OURS)), reason="oracle/_ref fastcar binaries not built")
def test_fastcar_same_output_as_the_reference(tmp_path):
    _compare(OURS, tmp_path)


@pytest.mark.skipif(not (os.path.isdir(os.path.join(REF_ROOT, "src", "fastcar")) and os.path.exists(REF)),
                    reason="reference sources / oracle/_ref not present")
@pytest.mark.timeout(600)
def test_fastcar_host_logic_on_cpu(tmp_path):
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "oracle"), "fastcar", "fastcar_stub", "REF=" + REF_ROOT])
    _compare(STUB, tmp_path)
