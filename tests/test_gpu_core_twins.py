"""Single-call twins of agents/cppmodule/core.cpp:20-26 on the GPU vs (a) the golden vectors produced by the
reference's own compiled core.cpp and (b) the C oracle on fresh seeded arenas.  Integer outputs (traces, visits,
unique lists, reachable sets) must be identical; value/variance are compared BIT-exactly as well, because the device
code reproduces the reference's float/double operation order (tolerance allowed by north_star: 1e-5)."""
import os

import numpy as np
import pytest

from arena_gen import make_arena, near_tie_arena
import gpu_twins as G

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "core_golden.npz")


def _leaves_and_traces(o, a, n, low=1, rng_state=None):
    return o.select_trace_obs(1, a["child"], a["visit"], a["value"], a["variance"], a["score"], a["n2o"], low, rng_state=rng_state)


@pytest.mark.parametrize("seed", range(6))
def test_select_matches_oracle(gpu_lib, oracle, seed):
    a = make_arena(seed, M=2048, max_depth=4 + seed)
    want = _leaves_and_traces(oracle, a, 1)
    got = G.select_trace_obs(1, a["child"], a["visit"], a["value"], a["variance"], a["score"], a["n2o"], 1)
    assert np.array_equal(got, want)
    # several roots deeper in the tree
    for r in want[:-1]:
        w = oracle.select_trace_obs(int(r), a["child"], a["visit"], a["value"], a["variance"], a["score"], a["n2o"], 1)
        g = G.select_trace_obs(int(r), a["child"], a["visit"], a["value"], a["variance"], a["score"], a["n2o"], 1)
        assert np.array_equal(g, w)


@pytest.mark.parametrize("seed", range(4))
def test_select_with_low_and_injected_rng(gpu_lib, oracle, seed):
    a = make_arena(100 + seed, M=2048, max_depth=6, unvisited=0.3)
    for low in (1, 5):
        s1 = np.array([0x1234567 + seed], np.uint32)
        s2 = s1.copy()
        want = oracle.select_trace_obs(1, a["child"], a["visit"], a["value"], a["variance"], a["score"], a["n2o"], low, rng_state=s1)
        got = G.select_trace_obs(1, a["child"], a["visit"], a["value"], a["variance"], a["score"], a["n2o"], low, rng_state=s2)
        assert np.array_equal(got, want) and s1[0] == s2[0]


def test_select_near_ties(gpu_lib, oracle):
    for seed in range(40):
        a = near_tie_arena(seed)
        want = oracle.select_trace_obs(1, a["child"], a["visit"], a["value"], a["variance"], a["score"], a["n2o"], 1)
        got = G.select_trace_obs(1, a["child"], a["visit"], a["value"], a["variance"], a["score"], a["n2o"], 1)
        assert np.array_equal(got, want), seed


@pytest.mark.parametrize("seed", range(4))
def test_backup_bit_exact(gpu_lib, oracle, seed):
    a = make_arena(200 + seed, M=2048, max_depth=8, unvisited=0.1)
    tr = oracle.select_trace_obs(1, a["child"], a["visit"], a["value"], a["variance"], a["score"], a["n2o"], 0)
    rng = np.random.default_rng(seed)
    for rep in range(5):
        v, var, gamma = float(rng.uniform(0, 400)), float(rng.uniform(0, 200)), float(rng.choice([0.999, 0.99, 0.9]))
        cpu = {k: a[k].copy() for k in ("visit", "value", "variance")}
        dev = {k: a[k].copy() for k in ("visit", "value", "variance")}
        oracle.backup_trace_obs(tr, cpu["visit"], cpu["value"], cpu["variance"], a["n2o"], a["score"], v, var, gamma)
        G.backup_trace_obs(tr, dev["visit"], dev["value"], dev["variance"], a["n2o"], a["score"], v, var, gamma)
        for k in cpu:
            assert np.array_equal(cpu[k], dev[k]), k
        a.update(cpu)


@pytest.mark.parametrize("mixture,averaged", [(0, 1), (0, 0), (1, 1), (1, 0)])
def test_backup_lp_bit_exact(gpu_lib, oracle, mixture, averaged):
    for seed in range(4):
        a = make_arena(300 + seed, M=2048, max_depth=6, unvisited=0.25)
        rng = np.random.default_rng(seed)
        # a trace that stops one level above a leaf, so the last node has children to initialise
        full = oracle.select_trace_obs(1, a["child"], a["visit"], a["value"], a["variance"], a["score"], a["n2o"], 0)
        tr = full[:-1] if len(full) > 1 else full
        cn, co = oracle.get_unique_child_obs(int(tr[-1]), a["child"], a["score"], a["n2o"])
        end = (rng.random(len(a["visit"])) < 0.1).astype(np.uint8)
        v = rng.uniform(0, 100, len(cn)).astype(np.float32)
        var = rng.uniform(0.1, 1000, len(cn)).astype(np.float32)
        cpu = {k: a[k].copy() for k in ("visit", "value", "variance")}
        dev = {k: a[k].copy() for k in ("visit", "value", "variance")}
        oracle.backup_trace_obs_LP(tr, cpu["visit"], cpu["value"], cpu["variance"], a["n2o"], a["score"], end, cn, co, v, var, 0.999, mixture, averaged)
        G.backup_trace_obs_LP(tr, dev["visit"], dev["value"], dev["variance"], a["n2o"], a["score"], end, cn, co, v, var, 0.999, mixture, averaged)
        for k in cpu:
            assert np.array_equal(cpu[k], dev[k]), (k, seed)
        # terminal-leaf form (no children): core.h:368-371
        e = np.zeros(0, np.int32)
        f = np.zeros(0, np.float32)
        oracle.backup_trace_obs_LP(full, cpu["visit"], cpu["value"], cpu["variance"], a["n2o"], a["score"], end, e, e, f, f, 0.999, mixture, averaged)
        G.backup_trace_obs_LP(full, dev["visit"], dev["value"], dev["variance"], a["n2o"], a["score"], end, e, e, f, f, 0.999, mixture, averaged)
        for k in cpu:
            assert np.array_equal(cpu[k], dev[k]), (k, seed, "terminal")


def test_unique_and_reachability(gpu_lib, oracle):
    for seed in range(4):
        a = make_arena(400 + seed, M=1024, max_depth=5, p_dup=0.3)
        for n in range(1, min(a["n_nodes"], 200)):
            wc, wo = oracle.get_unique_child_obs(n, a["child"], a["score"], a["n2o"])
            gc, go = G.get_unique_child_obs(n, a["child"], a["score"], a["n2o"])
            assert np.array_equal(wc, gc) and np.array_equal(wo, go), n
        for r in (1, 2, 5):
            assert G.get_all_childs(r, a["child"]) == oracle.get_all_childs(r, a["child"])


@pytest.mark.skipif(not os.path.exists(GOLD), reason="golden vectors not generated")
def test_against_reference_golden(gpu_lib):
    """Outputs of the reference's own core.cpp (compiled unchanged into oracle/_ref by tests/golden/gen_golden.py)."""
    z = np.load(GOLD)
    n_cases = int(z["n_cases"])
    for i in range(n_cases):
        p = "c%d_" % i
        a = {k: z[p + k] for k in ("child", "visit", "value", "variance", "score", "n2o")}
        low = int(z[p + "low"])
        got = G.select_trace_obs(1, a["child"], a["visit"], a["value"], a["variance"], a["score"], a["n2o"], low)
        assert np.array_equal(got, z[p + "trace"]), i
        cn, co = G.get_unique_child_obs(int(z[p + "uniq_node"]), a["child"], a["score"], a["n2o"])
        assert np.array_equal(cn, z[p + "uniq_c"]) and np.array_equal(co, z[p + "uniq_o"])
        dev = {k: a[k].copy() for k in ("visit", "value", "variance")}
        G.backup_trace_obs(z[p + "trace"], dev["visit"], dev["value"], dev["variance"], a["n2o"], a["score"],
                           float(z[p + "bk_v"]), float(z[p + "bk_var"]), float(z[p + "gamma"]))
        for k in dev:
            assert np.array_equal(dev[k], z[p + "bk_" + k]), (i, k)
        for m, av in ((0, 1), (0, 0), (1, 1), (1, 0)):
            q = p + "lp%d%d_" % (m, av)
            dev = {k: a[k].copy() for k in ("visit", "value", "variance")}
            G.backup_trace_obs_LP(z[p + "lp_trace"], dev["visit"], dev["value"], dev["variance"], a["n2o"], a["score"], z[p + "lp_end"],
                                  z[p + "lp_c"], z[p + "lp_o"], z[p + "lp_v"], z[p + "lp_var"], float(z[p + "gamma"]), m, av)
            for k in dev:
                assert np.array_equal(dev[k], z[q + k]), (i, m, av, k)
        assert G.get_all_childs(1, a["child"]) == set(z[p + "reach"].tolist())
