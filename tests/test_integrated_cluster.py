"""GPU, end to end: the reference's own meshclust2 with ONE translation unit swapped (src/cluster/Trainer.cpp ->
integration/Trainer_b200.cpp, which calls the C ABI) and the update stage's two loops offered to the device as one batch per
pass (integration/patch_cluster_factory.py) must produce the same clusters as the unmodified reference binary -- with the
batched update stage (default) and with the per-center calls (MC2_NO_BATCH=1).
Both binaries are built in the build container by `make -C oracle ref integrated` into oracle/_ref/ (they travel to the GPU
box; the test is skipped where they are absent).  --threads 1 and a single FASTA make the reference deterministic
(SURVEY.md section 4), and training is the reference's own host code in both, so the weights are identical by construction."""
import os
import re
import subprocess

import pytest

from conftest import ROOT
from meshclust2_b200 import synth

REF = os.path.join(ROOT, "oracle", "_ref", "meshclust2")
OURS = os.path.join(ROOT, "oracle", "_ref", "meshclust2_b200")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(OURS)), reason="oracle/_ref binaries not built")]


def parse_clstr(path):
    clusters, cur = [], None
    for line in open(path):
        if line.startswith(">Cluster"):
            cur = set()
            clusters.append(cur)
        else:
            m = re.search(r">(\S+)", line)
            if m:
                cur.add(m.group(1).rstrip("."))
    return {frozenset(c) for c in clusters if c}


def run(binary, fasta, workdir, out, env=None, extra=()):
    os.makedirs(workdir, exist_ok=True)
    r = subprocess.run([binary, "--id", "0.9", "--sample", "800", "--num-templates", "120", "--threads", "1", fasta] +
                       list(extra) + ["--output", out], cwd=workdir, capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    # what Runner::find_k and the width detection of Runner::run print (CRunner.cpp:497-498, :93, :109-121)
    decisions = [l for l in r.stdout.splitlines() if l.startswith(("avg length", "Recommended K", "Largest count", "Using "))]
    return open(os.path.join(workdir, "weights.txt")).read(), decisions


def test_same_clusters_as_the_reference(tmp_path):
    seqs, tids = synth.make_set(800, 1000, 120, 0.08, seed=7)
    fasta = str(tmp_path / "in.fa")
    open(fasta, "w").write(synth.to_fasta(seqs, tids))
    w_ref, d_ref = run(REF, fasta, str(tmp_path / "ref"), str(tmp_path / "ref.clstr"))
    assert len(d_ref) == 4, d_ref
    c_ref = parse_clstr(str(tmp_path / "ref.clstr"))
    assert sum(len(c) for c in c_ref) == 800 and 50 < len(c_ref) < 800
    # batched: update stage as one device batch per pass; ingest: also FASTA text -> segments -> k-mer histograms on the device
    # (declined for small files by default); k1: the reference's reader, histograms through K1; percall: one device call per
    # center, reader and histograms from the reference
    for tag, env in (("batched", {}), ("ingest", {"MC2_K1_MIN_BASES": "0"}),
                     ("k1", {"MC2_K1_MIN_BASES": "0", "MC2_NO_DEVICE_READER": "1"}), ("percall", {"MC2_NO_BATCH": "1"})):
        w_our, d_our = run(OURS, fasta, str(tmp_path / tag), str(tmp_path / (tag + ".clstr")), env)
        assert d_ref == d_our, (tag, d_ref, d_our)   # average length, k, largest count, histogram width
        assert w_ref == w_our                   # same host training code, same seeds
        c_our = parse_clstr(str(tmp_path / (tag + ".clstr")))
        assert sum(len(c) for c in c_our) == 800
        assert c_ref == c_our, "%s: clusters differ: %d vs %d clusters, %d in common" % (tag, len(c_ref), len(c_our), len(c_ref & c_our))


def test_no_train_list_same_clusters(tmp_path):
    """--no-train-list: the extra files' points join after the classifier was trained and every id is re-assigned
    (src/cluster/CRunner.cpp:576-592); the device mirror must follow the new numbering"""
    seqs, tids = synth.make_set(600, 1000, 90, 0.08, seed=17)
    train, extra = str(tmp_path / "train.fa"), str(tmp_path / "extra.fa")
    open(train, "w").write(synth.to_fasta(seqs[:350], tids[:350]))
    open(extra, "w").write("".join(">x%d template_%d\n%s\n" % (i, tids[350 + i], s.decode()) for i, s in enumerate(seqs[350:])))
    lst = str(tmp_path / "list.txt")
    open(lst, "w").write(extra + "\n")
    args = ("--no-train-list", lst)
    w_ref, _ = run(REF, train, str(tmp_path / "ref"), str(tmp_path / "ref.clstr"), extra=args)
    c_ref = parse_clstr(str(tmp_path / "ref.clstr"))
    assert sum(len(c) for c in c_ref) == 600
    for tag, env in (("batched", {}), ("percall", {"MC2_NO_BATCH": "1"}), ("noprewarm", {"MC2_NO_PREWARM": "1"})):
        w_our, _ = run(OURS, train, str(tmp_path / tag), str(tmp_path / (tag + ".clstr")), env, extra=args)
        assert w_ref == w_our
        c_our = parse_clstr(str(tmp_path / (tag + ".clstr")))
        assert c_ref == c_our, "%s: clusters differ: %d vs %d clusters, %d in common" % (tag, len(c_ref), len(c_our), len(c_ref & c_our))
