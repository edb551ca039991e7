"""the online agent's replay-memory policies on the device memory — accumulation policies 0-3, weighted_trimming, random_trimming,
store_nodes' random drop (agents/cppmodule/agent.cpp:619-819) — against the reference's OWN compiled agent.cpp (oracle/_ref, unmodified)
run with online=True: tests/golden/agent_online_golden.npz holds every memory it handed to its train callback and the move at which it did
(tests/golden/gen_golden.py gen_agent_cpp_online).  A single-game engine takes the policy step wherever the reference's remove_nodes() ran
(after every simulation step / update_root in which a collection happened); moments of training, row order, values and the played actions
must be identical."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "agent_online_golden.npz")


@pytest.mark.skipif(not os.path.exists(GOLD), reason="golden vectors not generated")
@pytest.mark.parametrize("policy", [0, 1, 2, 3])
def test_accumulation_policy_matches_compiled_agent_cpp(gpu_lib, policy):
    import torch
    from tetris_mcts_b200 import replay
    from tetris_mcts_b200.engine import BatchedEngine
    z = np.load(GOLD)
    p = "p%d_" % policy
    cfg = {k[len(p) + 4:]: int(z[k]) for k in z.files if k.startswith(p + "cfg_")}
    eng = BatchedEngine(1, max_nodes=cfg["M"], mode="lp", eval_kind="synthetic", lp_end_from_obs=True, lp_var_gamma2=False)   # agent.cpp:538,558
    eng.replay_enable(min_visits=cfg["min_visit"], capacity=cfg["memory_size"])
    eng.replay_policy(policy, episodes_per_train=cfg["ept"], memory_growth_rate=cfg["growth"])
    eng.set_games(z[p + "start"][None, :].astype(np.uint32))
    buf = torch.empty((2 * cfg["memory_size"], 212), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    calls, episode, gcs = [], 0, eng.counters()["gcs"]

    def after_collections(mv):
        nonlocal gcs
        now = eng.counters()["gcs"]
        while gcs < now:                                  # one policy step per remove_nodes() of the reference
            gcs += 1
            train_now, n = eng.replay_policy_step(eng.counters()["games_finished"])
            if train_now:
                eng.replay_peek_into(buf.data_ptr(), n)
                calls.append((mv, buf[:n].cpu().numpy().copy()))
                eng.replay_policy_trained(eng.counters()["games_finished"])

    actions = z[p + "actions"]
    for mv in range(len(actions)):
        for _ in range(cfg["sims"]):
            eng.run_sims(1)
            after_collections(mv)
        stats, action = eng.get_stats()
        assert action[0] == actions[mv], "search diverged from the reference's C++ agent at move %d" % mv
        eng.env_step(None)
        eng.update_root(auto_reset=True)
        after_collections(mv)
    assert len(calls) == int(z[p + "n_calls"]) and [c[0] for c in calls] == list(z[p + "call_moves"]), ([c[0] for c in calls], list(z[p + "call_moves"]))
    for i, (mv, rows) in enumerate(calls):
        states, value, variance, visit = replay.rows_to_memory(rows)
        assert np.array_equal(states[:, 0].astype(np.int8), z[p + "t%d_state" % i]), (policy, i, "state")
        assert np.array_equal(value[:, 0], z[p + "t%d_value" % i]), (policy, i, "value")
        assert np.array_equal(variance[:, 0], z[p + "t%d_variance" % i]), (policy, i, "variance")
        assert np.array_equal(visit[:, 0], z[p + "t%d_visit" % i]), (policy, i, "visit")
    eng.close()
