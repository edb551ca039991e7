"""oracle/mcts_oracle.c drop_tree — the CPU mirror of the engine's overflow_reset policy (NOT reference behaviour: the reference dies when
the reachable set fills the arena, agents/agent.py:99 IndexError / agents/cppmodule/agent.cpp:227-231).  Properties checked here on the CPU:
with the flag off the oracle reports the overflow like before; with it on the agent keeps simulating, drops trees, re-roots at the live game,
and is deterministic.  The engine is compared against this rule on the GPU (tests/test_gpu_bench_config.py)."""
import numpy as np


def play(oracle, overflow_reset, moves=60, M=256, sims=30):
    ag = oracle.Agent(max_nodes=M, mode=0, gamma=0.999, low=1, eval_mode=0, overflow_reset=overflow_reset)
    gm = oracle.Game(1, 0, 0)
    ag.update_root(gm.record())
    acts = []
    for _ in range(moves):
        ag.mcts(sims)
        a, st = ag.get_action()
        acts.append((a, st.copy()))
        gm.play(a)
        ag.update_root(gm.record())
        if gm.end:
            gm.reset()
            ag.update_root(gm.record())
    return ag, gm, acts


def test_reference_behaviour_reports_overflow(oracle):
    import pytest
    with pytest.raises(RuntimeError):
        play(oracle, 0)


def test_overflow_reset_keeps_playing_and_is_deterministic(oracle):
    a1, g1, acts1 = play(oracle, 1)
    a2, g2, acts2 = play(oracle, 1)
    assert a1.counter(7) > 0 and a1.counter(3) >= a1.counter(7)          # every dropped tree follows a collection
    assert a1.counter(0) == 60 * 30
    assert all(x[0] == y[0] and np.array_equal(x[1], y[1]) for x, y in zip(acts1, acts2))
    assert np.array_equal(g1.record(), g2.record())
    ex = a1.export()
    assert np.array_equal(ex["game"][a1.root], g1.record())              # rooted at the live game
    assert a1.n_free >= 256 // 8 or a1.counter(7) > 0
