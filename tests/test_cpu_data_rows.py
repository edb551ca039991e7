"""the reference's `State` row (util/Data.py:14-26) as a numpy structured dtype in PyTables' (alphabetical) column order."""
import numpy as np


def test_state_row_layout_matches_the_reference_description():
    from tetris_mcts_b200 import data as D
    want = {"episode": ("<i4", ()), "board": ("|i1", (20, 10)), "policy": ("<f4", (7,)), "action": ("|i1", ()), "combo": ("<i4", ()),
            "lines": ("<i4", ()), "line_stats": ("<i4", (4,)), "score": ("<i4", ()), "child_stats": ("<f4", (3, 7)), "cycle": ("<i4", ()),
            "value": ("<f4", ()), "variance": ("<f4", ())}                      # util/Data.py:15-26
    assert list(D.STATE_DTYPE.names) == sorted(want)                            # PyTables lays an IsDescription out alphabetically
    for name, (dt, shape) in want.items():
        f = D.STATE_DTYPE.fields[name][0]
        assert f.base.str == dt and f.shape == shape, name
    assert D.STATE_DTYPE.itemsize == 1 + 200 + 84 + 4 + 4 + 4 + 16 + 4 + 28 + 4 + 4 + 4


def test_saver_writes_rows_and_round_trips(tmp_path):
    from tetris_mcts_b200 import data as D
    n = 7
    rng = np.random.default_rng(0)
    rows = D.rows_from_batch(np.arange(n), rng.integers(-1, 2, (n, 20, 10)), rng.random((n, 7), np.float32), rng.integers(0, 7, n), 0,
                             rng.integers(0, 9, n), rng.integers(0, 3, (n, 4)), rng.integers(0, 999, n), rng.random((n, 3, 7), np.float32), 3, 1.5, 2.5)
    s = D.DataSaver(str(tmp_path) + "/", "data", 3, chunksize=5)
    s.add_rows(rows[:4]); s.add_rows(rows[4:])
    s.close()
    if D.have_pytables():
        import tables
        with tables.open_file(str(tmp_path / "data3")) as f:
            got = f.root.State.read()
    else:
        got = np.concatenate([np.load(str(p)) for p in sorted(tmp_path.glob("data3.*.npy"))])
    assert got.dtype == D.STATE_DTYPE and np.array_equal(got, rows)
