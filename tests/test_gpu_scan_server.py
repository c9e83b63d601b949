"""GPU: the resident scan server behind mc2_get_close (csrc/pair_score.cu scan_server_kernel: request / answer through a
mapped page-locked mailbox, no launch per call) against the launch path of the same entry point (MC2_NO_SCAN_SERVER=1) and
the oracle: Trainer<T>::get_close, src/cluster/Trainer.cpp:23-71.  Same arithmetic on both paths: results must be identical."""
import os
import time

import numpy as np
import pytest

from conftest import weights_path, weights_text
from helpers import synth_hist
from oracle import port

pytestmark = pytest.mark.gpu


def _both(ctx, fn):
    os.environ.pop("MC2_NO_SCAN_SERVER", None)
    a = fn()
    os.environ["MC2_NO_SCAN_SERVER"] = "1"
    try:
        b = fn()
    finally:
        os.environ.pop("MC2_NO_SCAN_SERVER", None)
    return a, b


@pytest.mark.parametrize("k,eb", [(5, 1), (5, 2), (6, 1)])
def test_server_equals_launch_path_and_oracle(built_lib, ctx, k, eb):
    rng = np.random.default_rng(3 + k + eb)
    n = 900
    H = synth_hist(rng, n, k, eb, hi=9)
    ln = rng.integers(850, 1200, n).astype(np.uint64)
    mag = H.sum(axis=1, dtype=np.uint64)
    hs = ctx.hset_from_host(H, k, length=ln)
    gm = ctx.model_from_file(weights_path("weights_cfg1_id90"))
    om = port.Model.from_text(weights_text("weights_cfg1_id90"))
    for trial in range(40):
        q = int(rng.integers(0, n))
        m = int(rng.choice([1, 2, 7, 10, 11, 33, 150, 192, 200, 640]))
        cand = rng.integers(0, n, m).astype(np.uint64)
        if trial % 5 == 4:
            time.sleep(0.002)          # longer than the server's idle time: it has left and is started again
        a, b = _both(ctx, lambda: ctx.get_close(gm, hs, q, hs, cand=cand, cutoff=0.9))
        assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2] and np.array_equal(a[3], b[3])
        ob = port.get_close(om, H, mag, ln, q, cand, 0.9)
        assert a[0] == ob[0] and a[2] == ob[2] and np.array_equal(a[3], ob[3])
        if ob[0] >= 0:
            assert abs(a[1] - ob[1]) <= 1e-9
    # contiguous candidate range, and a stale query magnitude / length (quirk Q4) through get_close_as
    a, b = _both(ctx, lambda: ctx.get_close(gm, hs, 5, hs, cand_begin=100, n_cand=300, cutoff=0.9))
    assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[3], b[3])
    a, b = _both(ctx, lambda: ctx.get_close_as(gm, hs, 7, int(mag[7]) + 40, 1000, hs, cand=np.arange(50, 400, dtype=np.uint64), cutoff=0.9))
    assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2] and np.array_equal(a[3], b[3])


def test_server_large_lists_take_the_launch_path(built_lib, ctx):
    rng = np.random.default_rng(1)
    n = 6000
    H = synth_hist(rng, n, 5, 1, hi=9)
    ln = rng.integers(900, 1100, n).astype(np.uint64)
    hs = ctx.hset_from_host(H, 5, length=ln)
    gm = ctx.model_from_file(weights_path("weights_cfg1_id90"))
    for m in (192, 193, 2049, 5000):
        cand = rng.integers(0, n, m).astype(np.uint64)
        a, b = _both(ctx, lambda: ctx.get_close(gm, hs, 3, hs, cand=cand, cutoff=0.9))
        assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[3], b[3])


def test_server_follows_model_and_row_changes(built_lib, ctx):
    """a second model, and rows rewritten between calls (mc2_hset_set_row is stream-ordered): the server must see both"""
    rng = np.random.default_rng(2)
    n = 300
    H = synth_hist(rng, n, 5, 1, hi=9)
    ln = rng.integers(900, 1100, n).astype(np.uint64)
    hs = ctx.hset_from_host(H, 5, length=ln)
    g1 = ctx.model_from_file(weights_path("weights_cfg1_id90"))
    g2 = ctx.model_from_file(weights_path("weights_appendixD_id90"))
    cand = np.arange(0, 200, dtype=np.uint64)
    for gm in (g1, g2, g1):
        a, b = _both(ctx, lambda: ctx.get_close(gm, hs, 250, hs, cand=cand, cutoff=0.9))
        assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[3], b[3])
    before = ctx.get_close(g1, hs, 250, hs, cand=cand, cutoff=0.9)
    hs.set_row(250, hs, 10)
    a, b = _both(ctx, lambda: ctx.get_close(g1, hs, 250, hs, cand=cand, cutoff=0.9))
    assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[3], b[3])
    assert a[1] != before[1] or not np.array_equal(a[3], before[3])
