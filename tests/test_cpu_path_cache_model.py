"""The coherence rules of the engine's path cache (tetris_mcts_b200/csrc/search_dev.cuh "path cache"), checked on the CPU.

oracle/mcts_oracle.c carries a model of the cache next to its (reference-pinned) tree agent: per level of the last trace the node, its
de-duplicated children and the statistics the engine would keep for them.  The model applies the engine's rules — k_backup refreshes the natural
copy of every trace level, truncates at the shallowest copy a transposition made stale, the walk drops the entry of a leaf it expanded,
collections / re-rooting / tree drops invalidate — and at EVERY selection compares every entry it would serve with the arena.  `errors` counts
entries that differ: it must be zero, i.e. a walk served from the cache reads exactly the values the reference's walk reads."""
import numpy as np
import pytest


def play(oracle, M, sims, moves, seed, overflow_reset=1, net=False):
    g = oracle.Game(seed=seed)
    a = oracle.Agent(max_nodes=M, mode=0, gamma=0.999, low=1, eval_mode=1 if net else 0, weights=oracle.seeded_weights(0) if net else None,
                     overflow_reset=overflow_reset)
    a.pc_enable(True)
    a.update_root(g.record())
    for _ in range(moves):
        a.mcts(sims)
        act, _ = a.get_action()
        g.play(act)
        a.update_root(g.record())
        if g.end:
            g.reset()
            a.update_root(g.record())
        if a.n_free < M * 5 // 32:
            a.remove_nodes()
    st = a.pc_stats()
    st["gcs"], st["resets"] = a.counter(3), a.counter(7)
    a.close()
    return st


@pytest.mark.parametrize("M,sims,moves,seed", [(16384, 300, 24, 123), (2048, 200, 60, 7), (1500, 60, 150, 5), (2048, 200, 60, 99)])
def test_cached_levels_always_equal_the_arena(oracle, M, sims, moves, seed):
    st = play(oracle, M, sims, moves, seed)
    assert st["errors"] == 0, st
    assert st["sims"] == sims * moves
    assert st["shared"] > 0.5 * st["levels"], st          # and the cache is worth having: most levels are served from it
    if M <= 2048:
        assert st["gcs"] > 0                              # collections (also in the middle of an expansion) happened on the way


def test_truncation_is_exercised(oracle):
    """Transpositions inside the path do occur (the same observation under two nodes of the path): the runs above must have hit the rule."""
    st = play(oracle, 2048, 200, 60, 99)
    assert st["irregular"] > 0 and st["errors"] == 0


def test_with_the_value_network_at_the_bench_arena_size(oracle):
    """The tree shape of the benchmarked configuration (value network as evaluator, 16384 slots, deep narrow trees): the cache serves ~90 % of the
    levels and every served value equals the arena's."""
    st = play(oracle, 16384, 250, 6, 123, net=True)
    assert st["errors"] == 0, st
    assert st["shared"] > 0.85 * st["levels"], st
