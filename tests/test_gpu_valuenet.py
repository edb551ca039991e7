"""Value network forward (model/model_vv.py:13-52, :210-217) on the GPU vs the C oracle and vs golden outputs of
the reference's own Model_VV (torch CPU).  Tolerance from north_star: 1e-5 (relative, fp32)."""
import os

import numpy as np
import pytest

from arena_gen import boards

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "valuenet_golden.npz")
RTOL, ATOL = 1e-5, 1e-5


def _boards_unused(n, seed):
    """tools/test.py:23-28 style boards: random {0,1}, top rows cleared, four -1 cells; plus real game positions."""
    rng = np.random.default_rng(seed)
    b = (rng.random((n, 20, 10)) < 0.45).astype(np.int8)
    for i in range(n):
        b[i, :rng.integers(2, 12)] = 0
        r, c = rng.integers(0, 3), rng.integers(0, 8)
        b[i, r:r + 2, c:c + 2] = -1
    return b


@pytest.mark.parametrize("kind", ["net"])
def test_matches_oracle(gpu_lib, oracle, kind):
    from tetris_mcts_b200.engine import BatchedEngine
    w = oracle.seeded_weights(0)
    eng = BatchedEngine(1, max_nodes=64, eval_kind=kind, weights=w)
    for n in (1, 3, 4, 7, 64, 250, 1000):
        s = boards(n, n)
        v, var = eng.valuenet(s)
        wv, wvar = oracle.valuenet_forward(w, s)
        assert np.allclose(v, wv, rtol=RTOL, atol=ATOL), np.abs(v - wv).max()
        assert np.allclose(var, wvar, rtol=RTOL, atol=ATOL), np.abs(var - wvar).max()
    eng.close()


def test_batch_position_independent(gpu_lib, oracle):
    """The same board must give the same bits wherever it sits in the batch (search determinism relies on it)."""
    from tetris_mcts_b200.engine import BatchedEngine
    eng = BatchedEngine(1, max_nodes=64, eval_kind="net", weights=oracle.seeded_weights(1))
    s = boards(300, 9)
    v1, r1 = eng.valuenet(s)
    perm = np.random.default_rng(0).permutation(300)
    v2, r2 = eng.valuenet(s[perm])
    assert np.array_equal(v1[perm], v2) and np.array_equal(r1[perm], r2)
    eng.close()


@pytest.mark.skipif(not os.path.exists(GOLD), reason="golden vectors not generated")
def test_matches_reference_golden(gpu_lib, oracle):
    from tetris_mcts_b200.engine import BatchedEngine
    z = np.load(GOLD)
    for seed in z["seeds"]:
        w = oracle.seeded_weights(int(seed))
        eng = BatchedEngine(1, max_nodes=64, eval_kind="net", weights=w)
        v, var = eng.valuenet(z["states"])
        assert np.allclose(v, z["v_%d" % seed], rtol=RTOL, atol=ATOL), np.abs(v - z["v_%d" % seed]).max()
        assert np.allclose(var, z["var_%d" % seed], rtol=RTOL, atol=ATOL), np.abs(var - z["var_%d" % seed]).max()
        eng.close()
