"""Board step parity: CUDA bitboard step (tetris_dev.cuh) vs the CPU restatement of SPEC_PYTETRIS.md.
PARITY UNPINNED by the reference (pyTetris source is absent): bit-exact here means vs oracle/tetris_oracle.c."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_fresh_games_match_oracle(gpu_lib, oracle):
    from tetris_mcts_b200 import pyTetris as PT
    for args in ((1, 0, 0), (1, 1, 1), (3, 0, 0), (2, 1, 0)):
        seeds = np.arange(1, 2001, dtype=np.uint32) * 7919
        got = PT.new_games(len(seeds), args, seeds)
        want = np.stack([oracle.Game(*args, seed=int(s)).record() for s in seeds])
        assert np.array_equal(got, want)
    assert np.array_equal(PT.new_games(1, (1, 0, 0), None)[0], oracle.Game(1, 0, 0).record())


@pytest.mark.parametrize("args", [(1, 0, 0), (1, 0, 1), (2, 1, 0), (4, 0, 0)])
def test_random_walk_bit_exact(gpu_lib, oracle, args):
    """>= 10^6 (state, action) pairs: 4096 games x 300 steps, each step compared record for record."""
    from tetris_mcts_b200 import pyTetris as PT
    n, steps = 4096, 300
    rng = np.random.default_rng(5)
    seeds = rng.integers(1, 2**32 - 1, n, dtype=np.uint64).astype(np.uint32)
    dev = PT.new_games(n, args, seeds)
    cpu = dev.copy()
    ended_total = 0
    for t in range(steps):
        actions = rng.integers(0, 8, n).astype(np.int32)     # 7 is out of range on purpose (VanillaC.py:7 randint(0,7))
        dev = PT.step_games(dev, actions)
        cpu = oracle.play_records(cpu, actions)
        assert np.array_equal(dev, cpu), "step %d" % t
        ended = (cpu[:, 10] >> 21) & 1
        if ended.any():    # SPEC §4 reset() on both sides
            idx = np.nonzero(ended)[0]
            ended_total += len(idx)
            sub = np.ascontiguousarray(dev[idx])
            gpu_lib.check(gpu_lib.lib().b200_tetris_new(gpu_lib.ptr(sub), len(idx), args[0], args[1], args[2], None, 1))
            dev[idx] = sub
            for i in idx:
                g = oracle.Game(record=cpu[i])
                g.reset()
                cpu[i] = g.record()
            assert np.array_equal(dev, cpu)
    assert ended_total > 0


def test_line_clears_and_scoring(gpu_lib, oracle):
    """Greedy-ish play that does clear lines, so lock / clear / combo / line_stats paths are compared too."""
    from tetris_mcts_b200 import pyTetris as PT
    n = 512
    rng = np.random.default_rng(11)
    dev = PT.new_games(n, (1, 0, 0), np.arange(100, 100 + n, dtype=np.uint32))
    cpu = dev.copy()
    # a scripted policy: shove pieces to alternating sides then hard drop -> fills rows
    lines_seen = 0
    for t in range(1500):
        phase = (t // 6) % 4
        a = np.where(rng.random(n) < 0.7, [1, 2, 1, 2][phase], 5 if t % 6 == 5 else rng.integers(0, 7)).astype(np.int32)
        dev = PT.step_games(dev, a)
        cpu = oracle.play_records(cpu, a)
        assert np.array_equal(dev, cpu)
        lines_seen = max(lines_seen, int(cpu[:, 14].max()))
        ended = np.nonzero((cpu[:, 10] >> 21) & 1)[0]
        for i in ended:
            g = oracle.Game(record=cpu[i]); g.reset(); cpu[i] = g.record()
        if len(ended):
            sub = np.ascontiguousarray(dev[ended])
            gpu_lib.check(gpu_lib.lib().b200_tetris_new(gpu_lib.ptr(sub), len(ended), 1, 0, 0, None, 1))
            dev[ended] = sub
    assert lines_seen > 0, "the scripted policy never cleared a line; the test lost its point"


def test_observation_matches(gpu_lib, oracle):
    from tetris_mcts_b200 import pyTetris as PT
    rng = np.random.default_rng(3)
    recs = PT.new_games(256, (1, 0, 0), np.arange(1, 257, dtype=np.uint32))
    for t in range(60):
        recs = PT.step_games(recs, rng.integers(0, 7, 256).astype(np.int32))
    st = PT.states_of(recs)
    for i in range(256):
        assert np.array_equal(st[i], oracle.Game(record=recs[i]).state())


def test_tetris_class_surface(gpu_lib, oracle):
    """The pyTetris.Tetris handle used by play.py (ctor, play, reset, copy_from, hash/eq, attributes)."""
    from tetris_mcts_b200.pyTetris import Tetris
    g = Tetris((20, 10), 1, 0, 0)
    o = oracle.Game(1, 0, 0)
    h = Tetris((20, 10), 1, 0, 0)
    assert g == h and hash(g) == hash(h)
    for a in [1, 1, 3, 5, 2, 6, 0, 5, 5, 4, 5]:
        g.play(a); o.play(a)
        assert np.array_equal(g.get_record(), o.record())
        assert np.array_equal(g.getState(), o.state())
    assert g != h
    h.copy_from(g)
    assert g == h and g.score == o.score and g.end == o.end and g.line_clears == o.line_clears
    g.reset(); o.reset()
    assert np.array_equal(g.get_record(), o.record())
    assert g.line_stats.shape == (4,) and g.combo == 0
