"""The whole per-move simulation loop on the GPU (k_select_expand -> evaluator -> k_backup, tree reuse, garbage
collection, root statistics) against the C oracle agent, which is itself pinned to the reference's UNMODIFIED
agents/agent.py + ValueSimLP.py + compiled core.cpp (tests/golden/gen_golden.py, tests/test_oracle_pins.py).
Everything is compared exactly: action, stats[3,7], and the full arena in the reference's array layout."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ARGS = (1, 0, 0)
GOLD = os.path.join(os.path.dirname(__file__), "golden", "agent_golden.npz")


def search_seed(seed, g):
    s = (seed + 0x9E3779B9 * (g + 1)) & 0xffffffff
    return s or 0x2545F491


def run_pair(oracle, mode, n, M, sims, moves, seed=123, eval_kind="synthetic", weights=None, engine_kw=None, agent_kw=None,
             eval_cb=None, check_arena=True):
    from tetris_mcts_b200 import pyTetris as PT
    from tetris_mcts_b200.engine import BatchedEngine
    mode_id = {"lp": 0, "single": 1, "vanilla": 2}[mode]
    recs = PT.new_games(n, ARGS, np.arange(seed, seed + n, dtype=np.uint32))
    eng = BatchedEngine(n, max_nodes=M, mode=mode, eval_kind=eval_kind, weights=weights, seed=seed, **(engine_kw or {}))
    eng.set_games(recs)
    kw = dict(max_nodes=M, mode=mode_id, gamma=0.99 if mode == "vanilla" else 0.999, low=5 if mode == "vanilla" else 1,
              eval_mode=0 if eval_cb is None else 2, eval_cb=eval_cb)
    kw.update(agent_kw or {})
    agents = [oracle.Agent(search_seed=search_seed(seed, g), **kw) for g in range(n)]
    games = [oracle.Game(record=recs[g]) for g in range(n)]
    for g in range(n):
        agents[g].update_root(games[g].record())
    gcs = 0
    for mv in range(moves):
        eng.run_sims(sims)
        stats, action = eng.get_stats()
        for g in range(n):
            agents[g].mcts(sims)
            a, st = agents[g].get_action()
            assert np.array_equal(st, stats[g]), "move %d game %d\n%s\n%s" % (mv, g, st, stats[g])
            assert a == action[g]
            games[g].play(a)
            agents[g].update_root(games[g].record())
            if games[g].end:
                games[g].reset()
                agents[g].update_root(games[g].record())
        eng.env_step(None)
        eng.update_root(auto_reset=True)
        assert np.array_equal(eng.get_games(), np.stack([gm.record() for gm in games])), "live games differ after move %d" % mv
    c = eng.counters()
    if check_arena:
        for g in range(min(n, 8)):
            ex = eng.export_game(g)
            want = agents[g].export()
            assert ex["root"] == agents[g].root
            for k in ("child", "n2o", "episode", "score", "visit", "value", "variance", "obs_end", "obs_key", "game"):
                assert np.array_equal(ex[k], want[k]), (g, k)
    assert c["sims"] == n * sims * moves
    assert c["gcs"] == sum(a.counter(3) for a in agents)
    assert c["expansions"] == sum(a.counter(1) for a in agents)
    eng.close()
    return c


def test_lp_synthetic_exact(gpu_lib, oracle):
    c = run_pair(oracle, "lp", n=24, M=4096, sims=60, moves=12)
    assert c["gcs"] == 0


def test_lp_with_garbage_collection(gpu_lib, oracle):
    """Small arenas force remove_nodes (agents/agent.py:206-257) in the middle of expansions, several times per game."""
    c = run_pair(oracle, "lp", n=12, M=1500, sims=12, moves=100)
    assert c["gcs"] >= 12


def test_lp_long_games_until_game_over(gpu_lib, oracle):
    c = run_pair(oracle, "lp", n=8, M=1500, sims=12, moves=140)
    assert c["games_finished"] >= 1


def test_lp_agent_cpp_variants(gpu_lib, oracle):
    """end taken from the observation (agent.cpp:538) and variance averaged without gamma^2 (agent.cpp:558)."""
    run_pair(oracle, "lp", n=8, M=2048, sims=40, moves=8, engine_kw=dict(lp_end_from_obs=True, lp_var_gamma2=False),
             agent_kw=dict(lp_end_from_obs=1, lp_var_gamma2=0))


def test_single_eval_mode_exact(gpu_lib, oracle):
    """ValueSim.py:52-94 (leaf itself evaluated; check_low draws from the injected RNG stream)."""
    run_pair(oracle, "single", n=16, M=2048, sims=60, moves=10)


def test_vanilla_rollouts_exact(gpu_lib, oracle):
    """Vanilla.py:17-64: low=5, gamma=0.99, random playouts from the shared xorshift stream, variance 1e3."""
    c = run_pair(oracle, "vanilla", n=16, M=4096, sims=80, moves=6)
    assert c["rollout_steps"] > 0


def test_no_stale_pop_variant(gpu_lib, oracle):
    run_pair(oracle, "lp", n=8, M=1500, sims=12, moves=100, engine_kw=dict(stale_pop=False), agent_kw=dict(stale_pop=0))


@pytest.mark.parametrize("kind", ["net", "net_tc"])
def test_real_network_search_is_exact_given_the_same_evaluator(gpu_lib, oracle, kind):
    """With the value network as evaluator the search must still be trace-for-trace identical when the oracle agent is
    fed the SAME network outputs (the LP path is deterministic given evaluator and piece sequence)."""
    from tetris_mcts_b200.engine import BatchedEngine
    w = oracle.seeded_weights(0)
    side = BatchedEngine(1, max_nodes=64, eval_kind=kind, weights=w)

    def cb(states):
        v, var = side.valuenet(states)
        return v, var

    run_pair(oracle, "lp", n=3, M=2048, sims=25, moves=4, eval_kind=kind, weights=w, eval_cb=cb)
    side.close()


def test_arena_overflow_is_reported_not_ub(gpu_lib):
    from tetris_mcts_b200 import pyTetris as PT
    from tetris_mcts_b200.engine import BatchedEngine
    from tetris_mcts_b200._lib import B200Error
    eng = BatchedEngine(4, max_nodes=40, mode="lp", eval_kind="synthetic")
    eng.set_games(PT.new_games(4, ARGS, np.arange(1, 5, dtype=np.uint32)))
    with pytest.raises(B200Error) as ei:
        for _ in range(50):
            eng.run_sims(20)
            eng.get_stats()
            eng.env_step(None)
            eng.update_root(True)
    assert ei.value.code == 3
    eng.close()


@pytest.mark.skipif(not os.path.exists(GOLD), reason="golden vectors not generated")
def test_against_reference_python_agent_golden(gpu_lib):
    """Per-move stats/actions recorded from the reference's own ValueSimLP (agents/ValueSimLP.py, agents/agent.py,
    compiled core.cpp) playing on the oracle env with the synthetic evaluator (tests/golden/gen_golden.py)."""
    from tetris_mcts_b200.engine import BatchedEngine
    z = np.load(GOLD)
    for case in range(int(z["n_cases"])):
        p = "a%d_" % case
        M, sims = int(z[p + "M"]), int(z[p + "sims"])
        eng = BatchedEngine(1, max_nodes=M, mode="lp", eval_kind="synthetic")
        eng.set_games(z[p + "start"].reshape(1, 20))
        for mv in range(len(z[p + "actions"])):
            eng.run_sims(sims)
            stats, action = eng.get_stats()
            assert np.array_equal(stats[0], z[p + "stats"][mv]), (case, mv)
            assert action[0] == z[p + "actions"][mv]
            eng.env_step(None)
            eng.update_root(True)
        ex = eng.export_game(0)
        for k in ("child", "score", "n2o", "visit", "value", "variance", "episode"):
            assert np.array_equal(ex[k], z[p + k]), (case, k)
        eng.close()


GOLD_MODES = os.path.join(os.path.dirname(__file__), "golden", "agent_modes_golden.npz")


@pytest.mark.skipif(not os.path.exists(GOLD_MODES), reason="golden vectors not generated")
def test_against_reference_valuesim_and_vanilla_golden(gpu_lib):
    """The other two mcts loops straight against the reference's own agents/ValueSim.py and agents/Vanilla.py (golden generated with
    their rand() / randint served from the xorshift stream the engine uses per game: oracle/rand_shim.c, gen_agent_modes):
    check_low picks, rollouts, leaf evaluation, collections — per-move statistics, actions and the final arrays, exact."""
    from tetris_mcts_b200.engine import BatchedEngine
    z = np.load(GOLD_MODES)
    for case in range(int(z["n_cases"])):
        p = "m%d_" % case
        mode = {1: "single", 2: "vanilla"}[int(z[p + "mode"])]
        seed = (int(z[p + "search_seed"]) - 0x9E3779B9) & 0xffffffff          # k_init_arena: srng[g] = seed + 0x9E3779B9 * (g + 1)
        eng = BatchedEngine(1, max_nodes=int(z[p + "M"]), mode=mode, eval_kind="synthetic", seed=seed)
        eng.set_games(z[p + "start"][None, :].astype(np.uint32))
        for mv, act in enumerate(z[p + "actions"]):
            eng.run_sims(int(z[p + "sims"]))
            stats, action = eng.get_stats()
            assert action[0] == act and np.array_equal(stats[0], z[p + "stats"][mv]), (case, mv)
            eng.env_step(None)
            eng.update_root(True)
        ex = eng.export_game(0)
        for k in ("child", "score", "n2o", "visit", "value", "variance", "episode"):
            assert np.array_equal(ex[k], z[p + k]), (case, k)
        assert ex["root"] == int(z[p + "root"])
        eng.close()


def test_full_size_properties(gpu_lib):
    """BASELINE config sizes are too large for the oracle; check size-independent invariants instead:
    every simulation adds exactly one visit to the root observation, child visits never exceed the root's,
    statuses stay clean, and two identical engines produce identical results (determinism)."""
    from tetris_mcts_b200 import pyTetris as PT
    from tetris_mcts_b200.engine import BatchedEngine
    n, sims = 16384, 40
    recs = PT.new_games(n, ARGS, np.arange(123, 123 + n, dtype=np.uint32))
    outs = []
    for rep in range(2):
        eng = BatchedEngine(n, max_nodes=1024, mode="lp", eval_kind="synthetic")
        eng.set_games(recs)
        eng.run_sims(sims)
        stats, action = eng.get_stats()
        ex = eng.export_game(n - 1)
        root_obs = ex["n2o"][ex["root"]]
        assert ex["visit"][root_obs] == sims
        assert (eng.status() == 0).all()
        assert stats[:, 0].sum(axis=1).max() <= sims * 7
        outs.append((stats.copy(), action.copy()))
        eng.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


GOLD_GC = os.path.join(os.path.dirname(__file__), "golden", "agent_gc_golden.npz")


@pytest.mark.skipif(not os.path.exists(GOLD_GC), reason="golden vectors not generated")
def test_explicit_remove_nodes_against_reference_python_agent_golden(gpu_lib):
    """The driver calling TreeAgent.remove_nodes() between moves whenever fewer than `headroom` slots are free
    (agents/agent.py:246-257 is a public method): b200_set_gc_headroom makes update_root do exactly that, batched.  Per-move
    statistics, actions and the final arrays against the reference's own ValueSimLP + agent.py (gen_agent_explicit_gc)."""
    from tetris_mcts_b200.engine import BatchedEngine
    z = np.load(GOLD_GC)
    for case in range(int(z["n_cases"])):
        p = "g%d_" % case
        M, sims, headroom = int(z[p + "M"]), int(z[p + "sims"]), int(z[p + "headroom"])
        eng = BatchedEngine(1, max_nodes=M, mode="lp", eval_kind="synthetic")
        eng.set_games(z[p + "start"][None, :].astype(np.uint32))
        eng.set_gc_headroom(headroom)                                   # the driver's policy starts after the first update_root
        gcs_at = []
        for mv, act in enumerate(z[p + "actions"]):
            g0 = eng.counters()["gcs"]
            eng.run_sims(sims)
            stats, action = eng.get_stats()
            assert action[0] == act and np.array_equal(stats[0], z[p + "stats"][mv]), (case, mv)
            g1 = eng.counters()["gcs"]
            eng.env_step(None)
            eng.update_root(True)
            if eng.counters()["gcs"] > g1:
                gcs_at.append(mv)
            assert g1 == g0, "no collection inside a move in these cases"
        ex = eng.export_game(0)
        for k in ("child", "score", "n2o", "visit", "value", "variance", "episode"):
            assert np.array_equal(ex[k], z[p + k]), (case, k)
        assert ex["root"] == int(z[p + "root"])
        assert gcs_at == list(z[p + "collected"]), (case, gcs_at)
        eng.close()


def test_batched_remove_nodes_many_games_exact(gpu_lib, oracle):
    """The same policy on many games at once (one k_gc launch for all of them) + an explicit collect-everything call, against
    oracle agents that make the same remove_nodes() calls."""
    from tetris_mcts_b200 import pyTetris as PT
    from tetris_mcts_b200.engine import BatchedEngine
    n, M, sims, moves, seed, headroom = 40, 2600, 24, 36, 555, 900
    recs = PT.new_games(n, ARGS, np.arange(seed, seed + n, dtype=np.uint32))
    eng = BatchedEngine(n, max_nodes=M, mode="lp", eval_kind="synthetic", seed=seed)
    eng.set_gc_headroom(headroom)
    eng.set_games(recs)
    agents = [oracle.Agent(max_nodes=M, mode=0, gamma=0.999, low=1, eval_mode=0, search_seed=search_seed(seed, g)) for g in range(n)]
    games = [oracle.Game(record=recs[g]) for g in range(n)]
    for g in range(n):
        agents[g].update_root(games[g].record())
    for mv in range(moves):
        actions, stats = eng.play_move(sims, auto_reset=True)
        for g in range(n):
            agents[g].mcts(sims)
            a, st = agents[g].get_action()
            assert a == actions[g] and np.array_equal(st, stats[g]), (mv, g)
            games[g].play(a)
            agents[g].update_root(games[g].record())
            if games[g].end:
                games[g].reset()
                agents[g].update_root(games[g].record())
            if agents[g].n_free < headroom:
                agents[g].remove_nodes()
        if mv == 10:                                            # TreeAgent.remove_nodes() on every game
            eng.remove_nodes()
            for g in range(n):
                agents[g].remove_nodes()
    assert eng.counters()["gcs"] == sum(ag.counter(3) for ag in agents) > n
    for g in (0, n // 2, n - 1):
        ex, want = eng.export_game(g), agents[g].export()
        for k in ("child", "score", "n2o", "visit", "value", "variance", "episode"):
            assert np.array_equal(ex[k], want[k]), (g, k)
    eng.close()


def test_overflow_reset_drops_trees_that_collection_cannot_shrink(gpu_lib):
    """overflow_reset (a policy beyond the reference, used by bench.py): when a collection recovers fewer than max_nodes/8 slots
    the whole thread block of k_gc clears the arena and the next kernel re-roots at the live game.  Invariants: no error status,
    trees are dropped (counter), every game keeps simulating, and the run is deterministic."""
    from tetris_mcts_b200 import pyTetris as PT
    from tetris_mcts_b200.engine import BatchedEngine
    n, M, sims, moves = 256, 512, 40, 30
    recs = PT.new_games(n, ARGS, np.arange(31, 31 + n, dtype=np.uint32))
    outs = []
    for rep in range(2):
        eng = BatchedEngine(n, max_nodes=M, mode="lp", eval_kind="synthetic", overflow_reset=True)
        eng.set_gc_headroom(200)
        eng.set_games(recs)
        acts = []
        for mv in range(moves):
            a, st = eng.play_move(sims, auto_reset=True)
            assert (st[:, 0].sum(axis=1) >= 1).mean() > 0.5             # (a tree dropped on the last simulations of a move has a bare root)
            acts.append(a.copy())
        c = eng.counters()
        assert (eng.status() == 0).all() and c["tree_resets"] > 0 and c["gcs"] > 0 and c["sims"] == n * sims * moves
        outs.append((np.stack(acts), eng.get_games().copy()))
        eng.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_graph_replay_equals_direct_launches(gpu_lib):
    """b200_run_sims replays one captured simulation step (CUDA graph) when phase timing is off and launches the
    kernels one by one when it is on; requests and counters are aggregated per CTA in k_select_expand.  Both paths must
    give the same statistics, actions, arenas and counters (including a garbage collection and the replay memory,
    which changes the captured kernel arguments after the first capture)."""
    from tetris_mcts_b200 import pyTetris as PT
    from tetris_mcts_b200.engine import BatchedEngine
    from tetris_mcts_b200.model.model_vv import init_weights
    n, M, sims, moves = 96, 1024, 25, 14
    recs = PT.new_games(n, ARGS, np.arange(7, 7 + n, dtype=np.uint32))
    for kind, w in (("synthetic", None), ("net_tc", init_weights(3))):
        res = []
        for timing in (False, True):
            eng = BatchedEngine(n, max_nodes=M, mode="lp", eval_kind=kind, weights=w, seed=5, overflow_reset=True)
            eng.set_timing(timing)
            eng.set_games(recs)
            out = []
            for mv in range(moves):
                if mv == 3:
                    eng.replay_enable(min_visits=2, capacity=50000)
                actions, stats = eng.play_move(sims, auto_reset=True)
                out.append((actions.copy(), stats.copy()))
            c = eng.counters()
            assert c["sims"] == n * sims * moves and c["gcs"] > 0
            res.append((out, c, eng.export_game(n // 2), eng.get_games().copy()))
            eng.close()
        (o0, c0, e0, g0), (o1, c1, e1, g1) = res
        for (a0, s0), (a1, s1) in zip(o0, o1):
            assert np.array_equal(a0, a1) and np.array_equal(s0, s1)
        for k in ("sims", "expansions", "eval_requests", "gcs", "trace_levels", "new_nodes"):
            assert c0[k] == c1[k], k
        for k in e0:
            assert np.array_equal(e0[k], e1[k]), k
        assert np.array_equal(g0, g1)


def test_engine_on_the_callers_stream(gpu_lib):
    """SURVEY 8(b).5 explicit cudaStream_t: the same moves on the engine's private stream, on a torch stream handed over with
    b200_engine_set_stream (switched in the middle of a game, graph re-captured), and back on a private stream give identical results;
    work the caller enqueues on ITS stream before a move (the upload of the games) is ordered before the engine's kernels."""
    import torch
    from tetris_mcts_b200 import pyTetris as PT
    from tetris_mcts_b200.engine import BatchedEngine
    n, M, sims, moves = 64, 1024, 20, 9
    recs = PT.new_games(n, ARGS, np.arange(11, 11 + n, dtype=np.uint32))
    res = []
    for external in (False, True):
        eng = BatchedEngine(n, max_nodes=M, mode="lp", eval_kind="synthetic", seed=9, overflow_reset=True)
        own = eng.get_stream()
        assert own != 0
        ts = torch.cuda.Stream()
        eng.set_games(recs)
        out = []
        for mv in range(moves):
            if external and mv == 2:
                eng.set_stream(ts)
                assert eng.get_stream() == ts.cuda_stream
            if external and mv == 6:
                eng.set_stream(None)
                assert eng.get_stream() not in (0, ts.cuda_stream)
            if external and mv == 4:       # caller-side work on the shared stream right before the move: a long kernel, then the engine's step
                with torch.cuda.stream(ts):
                    junk = torch.empty(1 << 26, device="cuda").normal_()
                    del junk
            actions, stats = eng.play_move(sims, auto_reset=True)
            out.append((actions.copy(), stats.copy()))
        res.append((out, eng.counters(), eng.export_game(3), eng.get_games().copy()))
        eng.close()
        ts.synchronize()                   # the caller's stream survives the engine
    (o0, c0, e0, g0), (o1, c1, e1, g1) = res
    for (a0, s0), (a1, s1) in zip(o0, o1):
        assert np.array_equal(a0, a1) and np.array_equal(s0, s1)
    for k in ("sims", "expansions", "eval_requests", "gcs", "trace_levels", "new_nodes"):
        assert c0[k] == c1[k], k
    for k in e0:
        assert np.array_equal(e0[k], e1[k]), k
    assert np.array_equal(g0, g1)


def test_path_cache_is_exact_against_the_oracle(gpu_lib, oracle):
    """b200_set_path_cache: the walk serves the levels it shares with the game's previous trace from the entries kept next to the trace
    (children's statistics refreshed by k_backup, truncated where a transposition made a copy stale, dropped by collections / re-rooting /
    tree drops).  Same oracle comparison as without it: actions, stats, arenas; with collections in the middle of expansions, explicit
    collections, dropped trees and games that end."""
    c = run_pair(oracle, "lp", n=24, M=4096, sims=60, moves=12, engine_kw=dict(path_cache=True))
    assert c["gcs"] == 0
    c = run_pair(oracle, "lp", n=12, M=1500, sims=12, moves=100, engine_kw=dict(path_cache=True))
    assert c["gcs"] >= 12
    c = run_pair(oracle, "lp", n=8, M=1500, sims=12, moves=140, engine_kw=dict(path_cache=True))
    assert c["games_finished"] >= 1
    run_pair(oracle, "lp", n=8, M=2048, sims=40, moves=8, engine_kw=dict(lp_end_from_obs=True, lp_var_gamma2=False, path_cache=True),
             agent_kw=dict(lp_end_from_obs=1, lp_var_gamma2=0))
    c = run_pair(oracle, "lp", n=16, M=65536, sims=300, moves=5, engine_kw=dict(path_cache=True), check_arena=False)   # the largest arena the cache serves, deep traces


def test_path_cache_on_equals_off_with_the_network_and_dropped_trees(gpu_lib):
    """Cache on against cache off on the production path (graph replay, net_tc, head-room collections, overflow_reset): every action, statistic,
    counter, exported arena and live game equal; and the cache must actually serve levels (it is not silently off)."""
    from tetris_mcts_b200 import pyTetris as PT
    from tetris_mcts_b200.engine import BatchedEngine
    from tetris_mcts_b200.model.model_vv import init_weights
    n, M, sims, moves = 256, 2048, 120, 10
    recs = PT.new_games(n, ARGS, np.arange(31, 31 + n, dtype=np.uint32))
    w = init_weights(1)
    res = []
    for pc in (False, True):
        eng = BatchedEngine(n, max_nodes=M, mode="lp", eval_kind="net_tc", weights=w, seed=5, overflow_reset=True, path_cache=pc)
        eng.set_gc_headroom(M * 5 // 32)
        eng.set_games(recs)
        out = []
        for mv in range(moves):
            actions, stats = eng.play_move(sims, auto_reset=True)
            out.append((actions.copy(), stats.copy()))
        c = eng.counters()
        res.append((out, c, [eng.export_game(g) for g in (0, n // 2, n - 1)], eng.get_games().copy()))
        eng.close()
    (o0, c0, e0, g0), (o1, c1, e1, g1) = res
    assert c0["gcs"] > 0 and c0["tree_resets"] > 0
    assert c0["cached_levels"] == 0 and c1["cached_levels"] > 0.6 * c1["trace_levels"], (c1["cached_levels"], c1["trace_levels"])
    for (a0, s0), (a1, s1) in zip(o0, o1):
        assert np.array_equal(a0, a1) and np.array_equal(s0, s1)
    for k in ("sims", "expansions", "eval_requests", "gcs", "trace_levels", "new_nodes", "tree_resets"):
        assert c0[k] == c1[k], k
    for x0, x1 in zip(e0, e1):
        for k in x0:
            assert np.array_equal(x0[k], x1[k]), k
    assert np.array_equal(g0, g1)


def test_path_cache_refused_where_it_does_not_apply(gpu_lib):
    from tetris_mcts_b200.engine import BatchedEngine
    for kw in (dict(mode="vanilla", max_nodes=1024, eval_kind="synthetic"), dict(mode="lp", max_nodes=65540, eval_kind="synthetic")):
        eng = BatchedEngine(2, **kw)
        with pytest.raises(gpu_lib.B200Error):
            eng.set_path_cache(True)
        eng.close()


def test_replay_memory_filled_at_garbage_collection(gpu_lib, oracle):
    """ValueSim.remove_nodes -> store_nodes(obs_available) (agents/ValueSim.py:101-159): the observations a collection frees,
    with visit >= min_visits_to_store and not end, as 212-byte rows.  The device stores them in arbitrary order, the
    reference in ascending index order: compared as sorted row sets."""
    import torch
    from tetris_mcts_b200 import pyTetris as PT
    from tetris_mcts_b200.engine import BatchedEngine
    n, M, sims, moves, seed, minv = 6, 1500, 12, 100, 123, 3
    recs = PT.new_games(n, ARGS, np.arange(seed, seed + n, dtype=np.uint32))
    eng = BatchedEngine(n, max_nodes=M, mode="lp", eval_kind="synthetic", seed=seed)
    eng.replay_enable(min_visits=minv, capacity=100000)
    eng.set_games(recs)
    agents = [oracle.Agent(max_nodes=M, mode=0, gamma=0.999, low=1, eval_mode=0, search_seed=search_seed(seed, g),
                           replay_min_visits=minv, replay_cap=100000) for g in range(n)]
    games = [oracle.Game(record=recs[g]) for g in range(n)]
    for g in range(n):
        agents[g].update_root(games[g].record())
    for mv in range(moves):
        actions, _ = eng.play_move(sims, auto_reset=True)
        for g in range(n):
            agents[g].mcts(sims)
            a, _st = agents[g].get_action()
            assert a == actions[g]
            games[g].play(a)
            agents[g].update_root(games[g].record())
            if games[g].end:
                games[g].reset()
                agents[g].update_root(games[g].record())
    want = np.concatenate([ag.replay() for ag in agents])
    assert len(want) > 50 and eng.counters()["gcs"] > 0
    buf = torch.zeros((100000, 212), dtype=torch.uint8, device="cuda")
    cnt = eng.replay_drain_into(buf.data_ptr(), 100000)
    got = buf[:cnt].cpu().numpy()
    assert cnt == len(want)
    key = lambda a: a[np.lexsort(a.T[::-1])]
    assert np.array_equal(key(got), key(want))
    assert eng.replay_drain_into(buf.data_ptr(), 100000) == 0          # drained: memory_index = 0 (ValueSim.py:183)
    from tetris_mcts_b200 import replay                                 # the same rows as ValueSim.memory arrays (ValueSim.py:25-30)
    states, values, variance, weights = replay.rows_to_memory(got)
    assert states.shape == (cnt, 1, 20, 10) and set(np.unique(states)) <= {-1.0, 0.0, 1.0} and (weights >= minv).all()
    eng.close()
