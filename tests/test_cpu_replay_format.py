"""The replay rows the engine emits re-laid out as the reference's ValueSim.memory arrays / dump.npz (agents/ValueSim.py:25-30,
122-159, 176-177): dtypes, shapes, keys and values, both directions."""
import numpy as np

from tetris_mcts_b200 import replay


def test_rows_to_reference_memory_and_dump(tmp_path):
    rng = np.random.default_rng(0)
    n = 37
    states = rng.integers(-1, 2, size=(n, 1, 20, 10)).astype(np.float32)
    values = rng.uniform(0, 90, (n, 1)).astype(np.float32)
    variance = rng.uniform(0.1, 900, (n, 1)).astype(np.float32)
    weights = rng.integers(25, 4000, (n, 1)).astype(np.float32)
    rows = replay.memory_to_rows(states, values, variance, weights)
    assert rows.shape == (n, 212) and rows.dtype == np.uint8
    mem = replay.rows_to_memory(rows)
    for got, want in zip(mem, (states, values, variance, weights)):       # ValueSim.py:25-30: float32, (N,1,20,10) / (N,1)
        assert got.dtype == np.float32 and got.shape == want.shape and np.array_equal(got, want)
    replay.dump(str(tmp_path / "dump"), rows)                             # ValueSim.py:177 -> ./data/dump.npz
    z = np.load(str(tmp_path / "dump.npz"))
    assert sorted(z.files) == ["states", "values", "variance", "weights"]
    assert np.array_equal(z["states"], states) and np.array_equal(z["weights"], weights)
    assert replay.rows_to_memory(np.zeros((0, 212), np.uint8))[0].shape == (0, 1, 20, 10)
