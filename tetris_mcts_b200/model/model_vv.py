"""Model_VV — host-side mirror of the reference value-network wrapper (model/model_vv.py:104-231, base model/model.py:39-255)
for the calls the agents make: Model_VV().load(); .training(False); .inference(batch) -> [v (k,1), var (k,1)].
The forward pass runs on the GPU through the C-ABI (b200_valuenet_forward); weights are the reference's state_dict
tensors (head.conv1.weight ... head.fc_out.bias, out_ubound, out_lbound) concatenated in that order."""
import os

import numpy as np

from .. import _lib as L

EXP_PATH = "./pytorch_model/"          # model/model.py:11
WEIGHT_KEYS = (("head.conv1.weight", (32, 1, 3, 3)), ("head.conv1.bias", (32,)), ("head.conv2.weight", (32, 32, 3, 3)),
               ("head.conv2.bias", (32,)), ("head.conv3.weight", (32, 32, 3, 3)), ("head.conv3.bias", (32,)),
               ("head.fc1.weight", (256, 1792)), ("head.fc1.bias", (256,)), ("head.fc_out.weight", (2, 256)),
               ("head.fc_out.bias", (2,)), ("out_ubound", (2,)), ("out_lbound", (2,)))


def init_weights(seed=0):
    """Random weights with the reference's default-init distribution (torch Conv2d/Linear: U(-1/sqrt(fan_in), +1/sqrt(fan_in))
    for weight and bias; out_ubound=[1e2,1e3], out_lbound=[0,1e-1], model_vv.py:45-46).  numpy PCG64 so that every box
    regenerates the same floats; no checkpoint of the current architecture ships with the reference."""
    rng = np.random.default_rng(seed)
    parts = []
    for shape, fan_in in (((32, 1, 3, 3), 9), ((32,), 9), ((32, 32, 3, 3), 288), ((32,), 288), ((32, 32, 3, 3), 288),
                          ((32,), 288), ((256, 1792), 1792), ((256,), 1792), ((2, 256), 256), ((2,), 256)):
        b = 1.0 / np.sqrt(fan_in)
        parts.append(rng.uniform(-b, b, size=shape).astype(np.float32).ravel())
    parts.append(np.array([1e2, 1e3], np.float32))
    parts.append(np.array([0.0, 1e-1], np.float32))
    return np.concatenate(parts)


def state_dict_to_weights(sd):
    """Flatten a reference checkpoint's model_state_dict (torch tensors or arrays) into the C-ABI weight vector."""
    parts = []
    for name, shape in WEIGHT_KEYS:
        t = sd[name]
        a = t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
        if tuple(a.shape) != shape:
            raise ValueError("%s: expected %s, checkpoint has %s (SURVEY §6: old checkpoints do not match the live Net)" % (name, shape, a.shape))
        parts.append(a.astype(np.float32).ravel())
    return np.concatenate(parts)


def load_checkpoint_weights(filename=EXP_PATH + "model_checkpoint"):
    """Model.load (model/model.py:163-174): the checkpoint's model_state_dict as the C-ABI weight vector, or None (after the
    reference's own message) when the file does not exist — the agent then keeps its default-initialised network."""
    if os.path.isfile(filename):
        import torch
        print("Loading model...", flush=True)
        ck = torch.load(filename, map_location="cpu")
        return state_dict_to_weights(ck["model_state_dict"])
    print("Checkpoint not found, using default model", flush=True)
    return None


class Model_VV:
    def __init__(self, device=0, seed=0, **kwargs):
        from ..engine import BatchedEngine
        self.weights = init_weights(seed)
        self._eng = BatchedEngine(1, max_nodes=64, eval_kind=kwargs.get("eval_kind", "net"), device=device)
        self._eng.load_weights(self.weights)

    def load(self, filename=EXP_PATH + "model_checkpoint"):       # model/model.py:163-174
        w = load_checkpoint_weights(filename)
        if w is not None:
            self.weights = w
            self._eng.load_weights(self.weights)

    def training(self, flag):                                     # inference only on this path
        if flag:
            raise NotImplementedError("training is outside the hot path (SURVEY §8f.2)")

    def inference(self, batch):                                   # model_vv.py:210-217
        b = np.asarray(batch)
        v, var = self._eng.valuenet(b.reshape(-1, 20, 10))
        return [v.reshape(-1, 1), var.reshape(-1, 1)]

    def close(self):
        self._eng.close()
