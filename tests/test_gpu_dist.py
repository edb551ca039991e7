"""Distributional cores on the GPU (agents/core_distributional.py twins) vs the reference's numba goldens and vs the
C oracle on fresh seeded arenas (including the low-visit random pick through the injected RNG stream)."""
import os

import numpy as np
import pytest

from arena_gen import make_dist_arena

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "dist_golden.npz")


def test_dist_twins_match_reference_golden(gpu_lib):
    from tetris_mcts_b200.agents import core_distributional as CD
    z = np.load(GOLD)
    for i in range(int(z["n_cases"])):
        p = "d%d_" % i
        assert np.allclose(CD.shift_distribution(z[p + "dist"], float(z[p + "x"]), 0, 5000), z[p + "shift"], rtol=1e-5, atol=1e-7)
        assert np.allclose(CD.mean_variance(z[p + "dist"], 0, 5000), z[p + "mv"], rtol=1e-5)
        tr = CD.select_trace_distributional(1, z[p + "child"], z[p + "node_stats"], z[p + "node_dist"], 0, 5000, low=0)
        assert np.array_equal(tr, z[p + "trace"]), i
        ns, nd = z[p + "node_stats"].copy(), z[p + "node_dist"].copy()
        CD.backup_trace_distributional(z[p + "trace"], ns, nd, float(z[p + "r"]), z[p + "dist"], 0, 5000)
        assert np.allclose(ns, z[p + "bk_stats"], rtol=1e-5, atol=1e-5)
        assert np.allclose(nd, z[p + "bk_dist"], rtol=1e-5, atol=1e-7)


def test_dist_twins_match_oracle_with_rng(gpu_lib, oracle):
    from tetris_mcts_b200.agents import core_distributional as CD
    rng = np.random.default_rng(1)
    for seed in range(6):
        a = make_dist_arena(50 + seed, M=512, max_depth=6, unvisited=0.25)
        for low in (0, 1, 5):
            CD.rng_state[0] = 777 + seed
            st = np.array([777 + seed], np.uint32)
            got = CD.select_trace_distributional(1, a["child"], a["node_stats"], a["node_dist"], 0, 5000, low=low)
            want = oracle.select_trace_distributional(1, a["child"], a["node_stats"], low, rng_state=st)
            assert np.array_equal(got, want) and CD.rng_state[0] == st[0]
        dist = rng.random(50).astype(np.float32)
        dist /= dist.sum()
        ns1, nd1, ns2, nd2 = a["node_stats"].copy(), a["node_dist"].copy(), a["node_stats"].copy(), a["node_dist"].copy()
        r = float(a["node_stats"][want[-1], 2] + 40)
        CD.backup_trace_distributional(want, ns1, nd1, r, dist, 0, 5000)
        oracle.backup_trace_distributional(want, ns2, nd2, r, dist, 0, 5000)
        assert np.array_equal(ns1, ns2) and np.array_equal(nd1, nd2)
