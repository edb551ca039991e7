"""the online agent's replay memory on the device.

(1) Policies and trimming — accumulation policies 0-3, weighted_trimming, random_trimming, store_nodes' order / full-stop / policy-0 random drop
(agents/cppmodule/agent.cpp:619-819) — against the reference's OWN code: tests/golden/replay_policy_golden.npz was recorded from
OnlineMCTSAgent::remove_nodes of the unmodified agent.cpp (compiled by inclusion into oracle/_ref/agent_probe, see oracle/agent_probe.cpp) driven
with scripted collections.  The engine gets the same collections (the rows a collection stores, in its order) and must report the same
memory_index after each, train at the same collections, and hand over bit-identical memories.

(2) Search through collections — the engine against the compiled agent.cpp run online for 420 moves per policy configuration
(tests/golden/agent_online_golden.npz): every action equal.  The MEMORY of those runs is not compared: TreeAgent::update_available
(agent.cpp:300-301) keeps the first n stale entries of its allocation history as "occupied", so which observations the reference frees (and
stores) depends on std::unordered_set iteration order — a defect a re-implementation cannot reproduce (oracle/agent_probe.cpp header)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "replay_policy_golden.npz")
GOLD_ONLINE = os.path.join(os.path.dirname(__file__), "golden", "agent_online_golden.npz")


def rows_of(states, value, variance, visit):
    n = len(value)
    rows = np.zeros((n, 212), np.uint8)
    rows[:, :200] = np.asarray(states, np.int8).reshape(n, 200).view(np.uint8)
    rows[:, 200:212] = np.ascontiguousarray(np.stack([value, variance, visit], axis=1).astype(np.float32)).view(np.uint8)
    return rows


@pytest.mark.skipif(not os.path.exists(GOLD), reason="golden vectors not generated")
@pytest.mark.parametrize("policy", [0, 1, 2, 3])
def test_policies_and_trimming_match_the_reference_remove_nodes(gpu_lib, policy):
    import torch
    from tetris_mcts_b200.engine import BatchedEngine
    z = np.load(GOLD)
    p = "p%d_" % policy
    cfg = {k[len(p) + 4:]: int(z[k]) for k in z.files if k.startswith(p + "cfg_")}
    eng = BatchedEngine(1, max_nodes=64, mode="lp", eval_kind="synthetic")
    eng.replay_enable(min_visits=cfg["min_visit"], capacity=cfg["memory_size"])
    eng.replay_policy(policy, episodes_per_train=cfg["ept"], memory_growth_rate=cfg["growth"])
    buf = torch.empty((2 * cfg["memory_size"], 212), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    call = 0
    for c in range(int(z[p + "n_collections"])):
        q = p + "c%d_" % c
        visit, end = z[q + "visit"], z[q + "end"]
        keep = (visit >= cfg["min_visit"]) & (end == 0)                      # store_nodes' filter (agent.cpp:797), k_gc's job in the engine
        eng.replay_append(rows_of(z[q + "states"][keep], z[q + "value"][keep], z[q + "variance"][keep], visit[keep].astype(np.float32)))
        episode = int(z[q + "episode"])
        train_now, n = eng.replay_policy_step(episode)
        assert int(train_now) == int(z[q + "trained"]), (policy, c)
        if train_now:
            eng.replay_peek_into(buf.data_ptr(), n)
            rows = buf[:n].cpu().numpy()
            t = p + "t%d_" % call
            assert n == len(z[t + "value"]), (policy, c, n)
            assert np.array_equal(rows[:, :200].view(np.int8).reshape(n, 20, 10), z[t + "state"]), (policy, c, "state")
            f = np.ascontiguousarray(rows[:, 200:212]).view(np.float32).reshape(n, 3)
            assert np.array_equal(f[:, 0], z[t + "value"]) and np.array_equal(f[:, 1], z[t + "variance"]) and np.array_equal(f[:, 2], z[t + "visit"]), (policy, c)
            eng.replay_policy_trained(episode)
            call += 1
            n = 0
        assert n == int(z[q + "memory_index"]), (policy, c, n, int(z[q + "memory_index"]))
    assert call == int(z[p + "n_calls"])
    eng.close()


@pytest.mark.skipif(not os.path.exists(GOLD_ONLINE), reason="golden vectors not generated")
@pytest.mark.parametrize("policy", [0, 3])
def test_search_equals_compiled_agent_cpp_through_collections(gpu_lib, policy):
    from tetris_mcts_b200.engine import BatchedEngine
    z = np.load(GOLD_ONLINE)
    p = "p%d_" % policy
    cfg = {k[len(p) + 4:]: int(z[k]) for k in z.files if k.startswith(p + "cfg_")}
    eng = BatchedEngine(1, max_nodes=cfg["M"], mode="lp", eval_kind="synthetic", lp_end_from_obs=True, lp_var_gamma2=False)   # agent.cpp:538,558
    eng.set_games(z[p + "start"][None, :].astype(np.uint32))
    actions = z[p + "actions"]
    for mv in range(len(actions)):
        a, _ = eng.play_move(cfg["sims"], auto_reset=True)
        assert a[0] == actions[mv], "search diverged from the reference's C++ agent at move %d" % mv
    assert eng.counters()["gcs"] >= 5
    eng.close()
