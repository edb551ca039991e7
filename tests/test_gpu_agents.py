"""The reference-facing Python surface driven exactly as play.py:75-181 drives it (constructor kwargs, update_root, play,
get_prob/get_stats/get_value_and_variance as util/Data.py:65-89 probes them, close), plus agents.cppmodule.core."""
import numpy as np
import pytest

from arena_gen import make_arena

pytestmark = pytest.mark.gpu


def play_loop(agent_cls, oracle, mode_id, sims, moves, gamma, low, **extra):
    from tetris_mcts_b200.pyTetris import Tetris
    env_args = ((20, 10), 1, 0, 0)                                  # play.py:75
    game = Tetris(*env_args)
    agent = agent_cls(sims=sims, env=Tetris, env_args=env_args, benchmark=False, online=False, min_visit=40, **extra)   # play.py:83-89
    agent.update_root(game)
    og = oracle.Game(1, 0, 0)
    oa = oracle.Agent(max_nodes=agent.max_nodes, mode=mode_id, gamma=gamma, low=low, eval_mode=0)
    oa.update_root(og.record())
    for mv in range(moves):
        action = agent.play()                                       # play.py:129
        oa.mcts(sims)
        a2, st2 = oa.get_action()
        assert action == a2 and np.array_equal(agent.get_stats(), st2)
        p = agent.get_prob()
        assert p.shape == (7,) and abs(p.sum() - 1) < 1e-5
        v, var = agent.get_value_and_variance()
        ex = oa.export()
        assert v == ex["value"][ex["n2o"][oa.root]]
        game.play(action)                                           # play.py:150
        og.play(action)
        agent.update_root(game)                                     # play.py:153
        oa.update_root(og.record())
        if game.end:
            game.reset(); og.reset()
            agent.update_root(game); oa.update_root(og.record())
    assert agent.root == oa.root
    assert np.array_equal(agent.arrays["child"], oa.export()["child"])
    assert np.array_equal(agent.node_to_obs, oa.export()["n2o"])
    assert agent.game_arr[agent.root] == game
    agent.close()                                                   # play.py:181


def test_valuesimlp_agent_like_play_py(gpu_lib, oracle):
    from tetris_mcts_b200.agents.ValueSimLP import ValueSimLP
    play_loop(ValueSimLP, oracle, 0, sims=40, moves=6, gamma=0.999, low=1, eval_kind="synthetic")


def test_valuesimlp_with_network_runs(gpu_lib):
    from tetris_mcts_b200.agents.ValueSimLP import ValueSimLP
    from tetris_mcts_b200.pyTetris import Tetris
    env_args = ((20, 10), 1, 0, 0)
    game = Tetris(*env_args)
    agent = ValueSimLP(sims=100, env=Tetris, env_args=env_args, benchmark=False, online=False, min_visit=40)   # BASELINE configs[0]
    agent.update_root(game)
    for _ in range(5):
        a = agent.play()
        assert 0 <= a < 7 and agent.get_stats()[0].sum() > 0
        game.play(a)
        agent.update_root(game)
    assert agent.counters()["sims"] == 500
    agent.close()


def test_core_module_signatures(gpu_lib, oracle):
    from tetris_mcts_b200.agents.cppmodule import core
    a = make_arena(7, M=512, max_depth=5, unvisited=0.2)
    tr = core.select_trace_obs(1, a["child"], a["visit"], a["value"], a["variance"], a["score"], a["n2o"], 0)
    assert np.array_equal(tr, oracle.select_trace_obs(1, a["child"], a["visit"], a["value"], a["variance"], a["score"], a["n2o"], 0))
    c, o = core.get_unique_child_obs(1, a["child"], a["score"], a["n2o"])
    wc, wo = oracle.get_unique_child_obs(1, a["child"], a["score"], a["n2o"])
    assert c == wc.tolist() and o == wo.tolist()
    dev = {k: a[k].copy() for k in ("visit", "value", "variance")}
    cpu = {k: a[k].copy() for k in ("visit", "value", "variance")}
    core.backup_trace_obs(tr, dev["visit"], dev["value"], dev["variance"], a["n2o"], a["score"], 12.5, 3.0, 0.999)
    oracle.backup_trace_obs(tr, cpu["visit"], cpu["value"], cpu["variance"], a["n2o"], a["score"], 12.5, 3.0, 0.999)
    for k in dev:
        assert np.array_equal(dev[k], cpu[k])
    assert core.get_all_childs(1, a["child"]) == oracle.get_all_childs(1, a["child"])
    with pytest.raises(TypeError):
        core.backup_trace_obs(tr, dev["visit"].astype(np.int64), dev["value"], dev["variance"], a["n2o"], a["score"], 1.0, 1.0, 0.9)


def test_dist_agent_class_like_play_py(gpu_lib, oracle):
    """play.py would do getattr(import_module('agents.DistValueSimOnline'), 'DistValueSimOnline') (play.py:81-82)."""
    from tetris_mcts_b200.agents import DistValueSimOnline as mod
    from tetris_mcts_b200.pyTetris import Tetris
    env_args = ((20, 10), 1, 0, 0)
    game = Tetris(*env_args)
    agent = getattr(mod, "DistValueSimOnline")(sims=80, env=Tetris, env_args=env_args, benchmark=True, online=False, min_visit=40)
    agent.update_root(game)
    for _ in range(3):
        a = agent.play()
        assert 0 <= a < 7
        game.play(a)
        agent.update_root(game)
    m, v = agent.get_value()
    assert 0 <= m <= 5000 and v >= 0
    agent.close()


def test_valuesim_loads_the_checkpoint_like_the_reference(gpu_lib, tmp_path, monkeypatch):
    """agents/ValueSim.py:42-44: Model().load() reads ./pytorch_model/model_checkpoint (model/model.py:163-174) when it exists.  A checkpoint
    on disk must change what the agent's network computes; without one the default-initialised weights are used."""
    import torch
    from tetris_mcts_b200.agents.ValueSimLP import ValueSimLP
    from tetris_mcts_b200.engine import BatchedEngine
    from tetris_mcts_b200.model.model_vv import WEIGHT_KEYS, init_weights
    from tetris_mcts_b200.pyTetris import Tetris
    env_args = ((20, 10), 1, 0, 0)
    monkeypatch.chdir(tmp_path)
    state = np.zeros((20, 10), np.int8)
    state[15:, :7] = 1
    state[2, 4:6] = -1; state[3, 4:6] = -1
    a0 = ValueSimLP(sims=10, env=Tetris, env_args=env_args, benchmark=True, online=False, min_visit=40)       # no checkpoint: default model
    v0 = a0.evaluate_state(state)
    a0.close()
    w = init_weights(5)
    sd, off = {}, 0
    for name, shape in WEIGHT_KEYS:
        n = int(np.prod(shape))
        sd[name] = torch.from_numpy(w[off:off + n].reshape(shape).copy())
        off += n
    (tmp_path / "pytorch_model").mkdir()
    torch.save({"model_state_dict": sd, "optimizer_state_dict": {}}, str(tmp_path / "pytorch_model" / "model_checkpoint"))   # model/model.py:143-160
    a1 = ValueSimLP(sims=10, env=Tetris, env_args=env_args, benchmark=True, online=False, min_visit=40)
    v1 = a1.evaluate_state(state)
    a1.close()
    ref = BatchedEngine(1, max_nodes=64, eval_kind="net_tc", weights=w)
    want = ref.valuenet(state[None])
    ref.close()
    assert (v1[0], v1[1]) == (want[0][0], want[1][0])
    assert (v0[0], v0[1]) != (v1[0], v1[1])
