"""CPU: the C-ABI library builds (nvcc cross-compiles), loads, and exports every symbol include/meshclust2_b200.h
declares; host-only entry points (input contract, weights parser) work; compute entry points fail loudly without a GPU."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, weights_path, weights_text
from oracle import port


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "meshclust2_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mc2_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(built_lib):
    L = built_lib.lib()
    syms = header_symbols()
    assert len(syms) >= 35
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(built_lib.SYMBOLS) == syms           # the binding's list and the header agree
    assert L.mc2_abi_version() == 1


def test_no_cpu_fallback(built_lib):
    if built_lib.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(built_lib.Mc2Error) as e:
        built_lib.Context(0)
    assert e.value.status == -2 and "no CPU fallback" in str(e.value)


def test_product_does_not_touch_oracle():
    """The product path must never import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "meshclust2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "mc2_oracle" not in src and "libmc2ref" not in src and "import oracle" not in src \
                    and "from oracle" not in src, f


def test_host_encode_matches_oracle(built_lib, golden, golden_seqs):
    off, soff = golden["text_off"], golden["seg_off"]
    for i, s in enumerate(golden_seqs):
        codes, segs, eff = built_lib.encode_dna(s)
        assert np.array_equal(codes, golden["codes"][off[i]:off[i + 1]]), i
        assert np.array_equal(segs, golden["segs"][soff[i]:soff[i + 1]]), i
        assert eff == golden["eff"][i]
    b = built_lib.encode_batch(golden_seqs, threads=2)
    assert np.array_equal(b["codes"], golden["codes"]) and np.array_equal(b["segs"], golden["segs"])
    assert np.array_equal(b["seg_off"].astype(np.int64), golden["seg_off"])
    assert np.array_equal(b["eff"].astype(np.int64), golden["eff"])
    rng = np.random.default_rng(0)
    for _ in range(200):
        s = bytes(rng.choice(list(b"ACGTNNNNacgtnRYKM"), size=int(rng.integers(1, 200))).tolist())
        o, c = port.encode(s), built_lib.encode_dna(s)
        assert np.array_equal(o[0], c[0]) and np.array_equal(o[1].reshape(-1, 2), c[1]) and o[2] == c[2]


def test_host_encode_rejects_invalid_letter(built_lib):
    with pytest.raises(built_lib.Mc2Error) as e:
        built_lib.encode_dna(b"ACGTACGTACGTACGTACGTACGTJACGT")
    assert e.value.status == -3


@pytest.mark.parametrize("wname", ["weights_cfg1_id90", "weights_appendixD_id90"])
def test_weights_parser_matches_oracle_parser(built_lib, wname):
    d, meta = built_lib.model_desc_from_file(weights_path(wname))
    m = port.Model.from_text(weights_text(wname))
    assert meta == dict(k=5, id=0.9, elem_bytes=1, mode=1)
    assert d.n_singles == len(m.singles) and d.n_combos == len(m.combos)
    for i, (f, lo, hi) in enumerate(m.singles):
        assert (d.single_flag[i], d.single_min[i], d.single_max[i]) == (f, lo, hi)
    for c, ((kind, _), idx) in enumerate(zip(m.combos, m.combo_indices())):
        assert d.combo_kind[c] == kind and list(d.combo_idx[c])[:d.combo_nidx[c]] == idx
    assert list(d.weight)[:len(m.weights)] == m.weights


def test_weights_parser_errors(built_lib, tmp_path):
    with pytest.raises(built_lib.Mc2Error) as e:
        built_lib.model_desc_from_file(str(tmp_path / "nope.txt"))
    assert e.value.status == -6
    p = tmp_path / "bad.txt"
    p.write_text("k: 5\nmode: 1\n")
    with pytest.raises(built_lib.Mc2Error):
        built_lib.model_desc_from_file(str(p))
    with pytest.raises(built_lib.Mc2Error):   # asks for a regression block the file does not have
        built_lib.model_desc_from_file(weights_path("weights_cfg1_id90"), which=1)
