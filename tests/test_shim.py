"""The reference-named C++ host classes (meshclust2_b200/host/mc2_shim.hpp) over the C ABI: compile everywhere,
run on the GPU against the reference-generated golden vectors."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, weights_path

SRC = os.path.join(ROOT, "tests", "cpp", "test_shim.cpp")
LIBDIR = os.path.join(ROOT, "meshclust2_b200", "lib")


def _build(tmp_path, built_lib):
    exe = str(tmp_path / "test_shim")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-o", exe, SRC, "-L" + LIBDIR, "-lmeshclust2_b200",
                           "-Wl,-rpath," + LIBDIR])
    return exe


def test_shim_compiles_and_links(tmp_path, built_lib):
    assert os.path.exists(_build(tmp_path, built_lib))


def _write_fixture(path, golden, golden_seqs, wname):
    # only plain-ACGT, one-segment sequences (the shim's string overloads feed raw text)
    keep = [i for i, s in enumerate(golden_seqs) if len(s) >= 200 and set(s) <= set(b"ACGT")][:40]
    H, ln, mag = golden["hist_k5_eb1"], golden["len_k5_eb1"], golden["mag_k5_eb1"]
    remap = {g: j for j, g in enumerate(keep)}
    ja, jb = golden["score_ia"], golden["score_ib"]
    sel = [j for j in range(len(ja)) if ja[j] in remap and jb[j] in remap][:120]
    with open(path, "w") as f:
        f.write("SEQS %d 5 1024\n" % len(keep))
        for i in keep:
            f.write(golden_seqs[i].decode() + "\n")
        f.write("HIST\n")
        for i in keep:
            f.write("%d %d %s\n" % (ln[i], mag[i], " ".join(map(str, H[i].tolist()))))
        S = golden[wname + "_cache"].shape[1]
        f.write("PAIRS %d %d\n" % (len(sel), S))
        for j in sel:
            f.write("%d %d %.17g %d %s\n" % (remap[ja[j]], remap[jb[j]], golden[wname + "_score"][j], golden[wname + "_close"][j],
                                             " ".join("%.17g" % v for v in golden[wname + "_cache"][j])))
        # get_close expectations restricted to the kept rows are recomputed with the oracle (test infrastructure)
        from oracle import port
        m = port.Model.from_text(open(weights_path(wname)).read())
        Hk, lk, mk = H[keep], ln[keep], mag[keep]
        cand = np.arange(len(keep))
        qs = list(range(0, len(keep), 7))
        f.write("GETCLOSE %d %d %s\n" % (len(qs), len(cand), " ".join(map(str, cand.tolist()))))
        for q in qs:
            best, bd, ismin, marks = port.get_close(m, Hk, mk, lk, q, cand, 0.9)
            f.write("%d %d %d %s\n" % (q, best, int(ismin), " ".join(map(str, marks.tolist()))))


@pytest.mark.gpu
def test_shim_against_golden(tmp_path, built_lib, golden, golden_seqs):
    exe = _build(tmp_path, built_lib)
    fx = str(tmp_path / "fixture.txt")
    _write_fixture(fx, golden, golden_seqs, "weights_cfg1_id90")
    r = subprocess.run([exe, fx, weights_path("weights_cfg1_id90")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr
