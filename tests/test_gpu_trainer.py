"""the value-network training step on the GPU (tetris_mcts_b200/csrc/trainer.cu) against goldens recorded from the
reference's OWN Model_VV (model/model_vv.py:104-231 + model/model.py:95-119 + model/yogi.py, torch CPU; tests/golden/gen_golden.py gen_train):
loss (GaussianLL, weighted std_mean), gradient norm, the gradients of the first step, and weights / Yogi state after the steps — rtol 1e-5
(the reference's fp32 arithmetic vs fp64-accumulating contractions rounded once)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "train_golden.npz")


def close(a, b, rtol=1e-5, atol=0.0, what=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = np.abs(b).max()
    err = np.abs(a - b).max()
    assert np.allclose(a, b, rtol=rtol, atol=atol + rtol * scale * 1e-2), "%s: max abs err %.3g (scale %.3g)" % (what, err, scale)


@pytest.mark.skipif(not os.path.exists(GOLD), reason="golden vectors not generated")
@pytest.mark.parametrize("tag,weighted,clip", [("w", True, 0.0), ("u", False, 0.0), ("c", True, 0.5)])
def test_training_steps_match_the_reference(gpu_lib, tag, weighted, clip):
    from tetris_mcts_b200.model.model_vv import init_weights
    from tetris_mcts_b200.model.trainer import Trainer
    z = np.load(GOLD)
    keep = z["keep_index"]
    batch = [z["states"], z["value"], z["variance"], z["weight"]]
    t = Trainer(init_weights(int(z["seed"])), max_batch=128)
    t.set_out_ubound(*z["ubound"])                                       # model_vv.py:227-231
    steps = z[tag + "_steps"]
    for it in range(len(steps)):
        r = t.step(batch, weighted=weighted, grad_clip=clip)
        assert abs(r["loss"] - steps[it, 0]) <= 1e-5 * abs(steps[it, 0]) + 1e-6, (it, r, steps[it])
        assert abs(r["loss_std"] - steps[it, 1]) <= 1e-5 * abs(steps[it, 1]) + 1e-6, (it, r, steps[it])
        assert abs(r["grad_norm"] - steps[it, 2]) <= 1e-5 * abs(steps[it, 2]) + 1e-6, (it, r, steps[it])
        if it == 0:
            close(t.grads()[keep], z[tag + "_grad0"], what="gradients of step 1")
    close(t.weights()[:478338][keep], z[tag + "_weights"], rtol=1e-5, what="weights")
    if tag == "w":
        m, v, step = t.state()
        assert step == len(steps)
        close(m[keep], z["w_exp_avg"], what="exp_avg")
        close(v[keep], z["w_exp_avg_sq"], what="exp_avg_sq")
    t.close()


@pytest.mark.skipif(not os.path.exists(GOLD), reason="golden vectors not generated")
def test_validation_loss_inference_and_device_rows(gpu_lib, tmp_path):
    """Model.compute_loss (model/model.py:52-83) and inference after three steps; the same step from 212-byte replay rows gathered on the
    device (the producer format of k_gc / the all-gather) gives the same numbers as the host-batch step."""
    import torch
    from tetris_mcts_b200 import replay
    from tetris_mcts_b200.model.model_vv import Model_VV, init_weights
    from tetris_mcts_b200.model.trainer import Trainer
    z = np.load(GOLD)
    batch = [z["states"], z["value"], z["variance"], z["weight"]]
    m = Model_VV(seed=int(z["seed"]))
    m._trainer_obj().set_out_ubound(*z["ubound"])
    for _ in range(3):
        m.train(batch, weighted=True)
    val = m.compute_loss([b[:40] for b in batch], weighted=True, chunksize=16)
    assert abs(val["loss"] - z["w_val"][0]) < 1e-5 * abs(z["w_val"][0]) and abs(val["loss_std"] - z["w_val"][1]) < 1e-5 * abs(z["w_val"][1])
    m._publish()
    v, var = m.inference(z["states"][:8, None])
    assert np.allclose(np.concatenate([v, var], 1), z["w_pred"], rtol=1e-5, atol=1e-6)
    # checkpoint round trip in the reference's file layout (model/model.py:143-174)
    ck = str(tmp_path / "model_checkpoint")
    m.save(ck, verbose=False)
    d = torch.load(ck, map_location="cpu", weights_only=False)
    assert list(d["model_state_dict"])[:2] == ["head.conv1.weight", "head.conv1.bias"] and d["optimizer_state_dict"]["state"][0]["step"] == 3
    m2 = Model_VV(seed=1)
    m2.load(ck)
    assert np.array_equal(m2.weights, m.weights)
    r_a, r_b = m.train(batch, weighted=True), m2.train(batch, weighted=True)
    assert r_a == r_b, "optimizer state was not restored"
    m.close(); m2.close()
    # device-side gather from replay rows
    w = init_weights(0)
    rows = replay.memory_to_rows(z["states"], z["value"], z["variance"], np.round(z["weight"] * 100))
    scale = 1.0 / float(np.round(z["weight"] * 100).mean())
    t1, t2 = Trainer(w, max_batch=128), Trainer(w, max_batch=128)
    idx = np.arange(len(rows), dtype=np.int32)[::-1].copy()
    wts = (np.round(z["weight"] * 100) * np.float32(scale)).astype(np.float32)
    r1 = t1.step([z["states"][idx], z["value"][idx], z["variance"][idx], wts[idx]], weighted=True)
    dev_rows = torch.from_numpy(rows).cuda()
    torch.cuda.synchronize()
    r2 = t2.step_rows_dev(dev_rows.data_ptr(), len(rows), idx, scale, weighted=True)
    assert r1 == r2 and np.array_equal(t1.weights(), t2.weights())
    t1.close(); t2.close()


def test_train_data_loop_lowers_the_loss_and_hot_swaps(gpu_lib, tmp_path, monkeypatch):
    """Model.train_data (model/model.py:176-249): split, batches, validation lines in the reference's log format (web/parseLog.py:61-66),
    early stopping with the best checkpoint re-loaded; the inference kernels then use the trained weights."""
    import io
    import re
    from tetris_mcts_b200 import pyTetris as PT
    from tetris_mcts_b200.model import model_vv as MV
    from tetris_mcts_b200.model.model_vv import Model_VV
    monkeypatch.chdir(tmp_path)
    log = io.StringIO()
    monkeypatch.setattr(MV, "perr", dict(file=log, flush=True))                  # the reference prints these lines to stderr (model/model.py:13)
    rng = np.random.default_rng(3)
    recs = PT.new_games(64, (1, 0, 0), np.arange(5, 69, dtype=np.uint32))
    states = []
    for _ in range(12):
        recs = PT.step_games(recs, rng.integers(0, 7, 64))
        states.append(PT.states_of(recs))
    states = np.concatenate(states)[:, None].astype(np.float32)
    n = len(states)
    filled = (states > 0).sum(axis=(1, 2, 3)).astype(np.float32)[:, None]
    data = [states, 2.0 * filled + 5, 10.0 + filled, rng.integers(25, 200, (n, 1)).astype(np.float32)]
    m = Model_VV(seed=2)
    before = m.compute_loss([d[-76:] for d in [data[0], data[1], data[2], data[3] / data[3].mean()]], weighted=True)["loss"]
    np.random.seed(0)
    m.train_data([d.copy() for d in data], batch_size=64, iters_per_val=25, max_iters=200)
    err = log.getvalue()
    train_re = r'Iteration:\s*(?P<iter>\d*)\s*training loss:\s*(?P<t_loss>\d*\.\d*)\s*validation loss:\s*(?P<v_loss>\d*\.\d*)±\s*(?P<v_loss_err>\d*\.\d*|nan)\s*gradient norm:\s*(?P<g_norm>\d*\.\d*)'
    assert re.search(r'Training data size:\s*(\d*)\s*Validation data size:\s*(\d*)', err) and len(re.findall(train_re, err)) >= 4
    after = m.compute_loss([d[-76:] for d in [data[0], data[1], data[2], data[3] / data[3].mean()]], weighted=True)["loss"]
    assert after < before - 0.5, (before, after)
    assert os.path.isfile("pytorch_model/model_checkpoint")
    v, var = m.inference(states[:4])
    t_pred = m._trainer_obj().loss([states[:4], data[1][:4], data[2][:4], None], weighted=False, want_pred=True)[2]
    assert np.allclose(np.concatenate([v, var], 1), t_pred, rtol=1e-5)
    m.close()
