"""the reference's log / status wire formats from a batched run.  The `Episode:` lines must match the regular expression
web/parseLog.py:58-60 parses them with, the realtime memmaps the dtypes and shapes web/parseLog.py:34-38 maps them with, and every
finished episode's numbers must equal the CPU oracle env replaying the same actions."""
import io
import re

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SCORE_RE = r'Episode:\s*(?P<episode>\d*)\s*Score:\s*(?P<score>\d*)\s*Lines Cleared:\s*(?P<lines>\d*)'      # web/parseLog.py:58-60


def test_finished_games_log_matches_oracle_env(gpu_lib, oracle):
    from tetris_mcts_b200 import pyTetris as PT
    from tetris_mcts_b200.engine import BatchedEngine
    n, seed = 64, 900
    recs = PT.new_games(n, (1, 0, 0), np.arange(seed, seed + n, dtype=np.uint32))
    eng = BatchedEngine(n, max_nodes=4096, mode="vanilla", eval_kind="synthetic", seed=seed, overflow_reset=True)
    eng.set_games(recs)
    games = [oracle.Game(record=recs[g]) for g in range(n)]
    want, got, episodes = [], [], [0] * n
    for mv in range(120):
        actions, _ = eng.play_move(20, auto_reset=True)
        for g in range(n):
            games[g].play(int(actions[g]))
            if games[g].end:
                episodes[g] += 1
                want.append((g, games[g].score, games[g].line_clears, episodes[g]))
                games[g].reset()
        rows = eng.finished_games()
        got += [tuple(int(x) for x in r) for r in rows]
    assert len(want) > 10 and got == want
    assert eng.counters()["games_finished"] == len(want)
    eng.close()


def test_play_batched_emits_the_reference_wire_formats(gpu_lib, tmp_path):
    from tetris_mcts_b200 import play_batched as PB
    out = io.StringIO()
    argv = ["--agent_type", "Vanilla", "--mcts_sims", "16", "--ngames", "12", "--n_parallel", "32", "--max_nodes", "2048", "--endless",
            "--realtime_status", "--status_dir", str(tmp_path / "tmp"), "--max_moves", "400"]
    PB.main(argv, out=out)
    lines = [ln for ln in out.getvalue().splitlines() if ln.startswith("Episode:")]
    assert len(lines) == 12
    for i, ln in enumerate(lines):
        m = re.search(SCORE_RE, ln)
        assert m and int(m.group("episode")) == i + 1 and m.group("score") != "" and m.group("lines") != ""
        assert ln == PB.EPISODE_FMT.format(i + 1, int(m.group("score")), int(m.group("lines")))
    # StatusParser (web/parseLog.py:34-38)
    board = np.memmap(str(tmp_path / "tmp" / "board"), mode='r', dtype=np.int8, shape=(20, 10))
    combo = np.memmap(str(tmp_path / "tmp" / "combo"), mode='r', dtype=np.int32, shape=(1,))
    lines_mm = np.memmap(str(tmp_path / "tmp" / "lines"), mode='r', dtype=np.int32, shape=(1,))
    score = np.memmap(str(tmp_path / "tmp" / "score"), mode='r', dtype=np.int32, shape=(1,))
    line_stats = np.memmap(str(tmp_path / "tmp" / "line_stats"), mode='r', dtype=np.int32, shape=(4,))
    assert set(np.unique(board)) <= {-1, 0, 1} and (board == -1).sum() == 4
    assert combo[0] >= 0 and score[0] >= 0 and lines_mm[0] == line_stats[0] + 2 * line_stats[1] + 3 * line_stats[2] + 4 * line_stats[3]


def test_play_batched_saves_state_rows(gpu_lib, tmp_path):
    """--save: one util/Data.py `State` row per game and move (play.py:131-132), written by tetris_mcts_b200.data.DataSaver."""
    from tetris_mcts_b200 import data as D
    from tetris_mcts_b200 import play_batched as PB
    out = io.StringIO()
    PB.main(["--agent_type", "Vanilla", "--mcts_sims", "8", "--ngames", "1000", "--n_parallel", "16", "--max_nodes", "1024", "--endless", "--max_moves", "6",
             "--save", "--save_dir", str(tmp_path) + "/", "--save_file", "data", "--cycle", "2"], out=out)
    if D.have_pytables():
        import tables
        with tables.open_file(str(tmp_path / "data2")) as f:
            rows = f.root.State.read()
    else:
        rows = np.concatenate([np.load(str(p)) for p in sorted(tmp_path.glob("data2.*.npy"))])
    assert len(rows) == 16 * 6 and rows.dtype == D.STATE_DTYPE
    assert set(np.unique(rows["board"])) <= {-1, 0, 1} and ((rows["board"] == -1).sum(axis=(1, 2)) == 4).all()
    assert np.allclose(rows["policy"].sum(axis=1), 1.0, atol=1e-5) and (rows["cycle"] == 2).all()
    assert ((rows["action"] >= 0) & (rows["action"] < 7)).all() and (rows["child_stats"][:, 0].sum(axis=1) > 0).all()
