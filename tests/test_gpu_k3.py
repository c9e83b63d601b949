"""GPU: K3 (cluster mean + closest member, mc2_mean_closest / mc2_closest) vs the oracle. The mean is bit-exact, the
distances too (same IEEE operations, no fused multiply-add on either side), so the arg-min is identical."""
import numpy as np
import pytest

from oracle import port

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,eb", [(5, 1), (3, 1), (1, 1), (5, 2), (4, 4), (3, 8), (6, 1)])
def test_mean_closest_vs_oracle(built_lib, ctx, k, eb):
    rng = np.random.default_rng(100 * k + eb)
    N = 4 ** k
    hi = {1: 255, 2: 3000, 4: 100000, 8: 100000}[eb]
    H = rng.integers(1, hi + 1, size=(300, N)).astype(port.DTYPES[eb])
    hs = ctx.hset_from_host(H, k)
    for n in (1, 2, 7, 130, 300):
        mem = rng.integers(0, 300, n)
        best, bd, mean, dist = ctx.mean_closest(hs, mem)
        ob, obd, omean, odist = port.mean_closest(H, mem)
        assert np.array_equal(mean, omean)
        assert np.array_equal(dist, odist)
        assert best == ob and bd == obd
        b2, bd2, d2 = ctx.closest(hs, mem, omean)          # Trainer::closest with a host-built mean
        assert b2 == ob and bd2 == obd and np.array_equal(d2, odist)


def test_first_minimum_wins_on_ties(built_lib, ctx):
    rng = np.random.default_rng(3)
    H = rng.integers(1, 9, size=(6, 1024)).astype(np.uint8)
    H[4] = H[1]                                            # duplicates: identical distances
    hs = ctx.hset_from_host(H, 5)
    mem = np.array([4, 1, 1, 4, 0])
    best, bd, mean, dist = ctx.mean_closest(hs, mem)
    ob = port.mean_closest(H, mem)
    assert best == ob[0] and dist[0] == dist[1] == dist[2]


def test_golden_distance_d(built_lib, ctx, golden):
    H = golden["hist_k5_eb1"]
    hs = ctx.hset_from_host(H, 5)
    C, ib = golden["distance_d_centers"], golden["pair_ib"]
    for j in range(40):
        _, _, d = ctx.closest(hs, [int(ib[j])], C[j])
        assert abs(d[0] - golden["distance_d_k5_eb1"][j]) <= 1e-12 * max(1.0, golden["distance_d_k5_eb1"][j])


def test_empty_member_list_is_an_error(built_lib, ctx):
    hs = ctx.hset_from_host(np.ones((2, 16), dtype=np.uint8), 2)
    with pytest.raises(built_lib.Mc2Error):
        ctx.mean_closest(hs, np.zeros(0, dtype=np.uint64))
