"""ref_net.py — TEST INFRASTRUCTURE ONLY.  torch (CPU, fp32) restatement of the reference value network module
(model/model_vv.py:13-52 Net: head = conv1/act/conv2/act/conv3/act/flatten/fc1/act/fc_out/sigmoid, then
`x * out_ubound + out_lbound`) and of Model_VV.inference (:210-217).  It exists because the reference's Python files
cannot travel to the GPU box: bench.py's reference arm drives the reference's own compiled agents/cppmodule/agent.cpp
(oracle/_ref/agent*.so) and needs the evaluator callback that agents/ValueSimC.py:17-42 passes (Model_VV.inference)."""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        act = nn.ReLU(inplace=True)
        self.head = nn.Sequential(OrderedDict([
            ("conv1", nn.Conv2d(1, 32, 3, 1)), ("act1", act), ("conv2", nn.Conv2d(32, 32, 3, 1)), ("act2", act),
            ("conv3", nn.Conv2d(32, 32, 3, 1)), ("act3", act), ("flatten", nn.Flatten()), ("fc1", nn.Linear(1792, 256)),
            ("fc_act1", act), ("fc_out", nn.Linear(256, 2)), ("act_out", nn.Sigmoid())]))
        self.out_ubound = nn.Parameter(torch.tensor([1e2, 1e3]), requires_grad=False)
        self.out_lbound = nn.Parameter(torch.tensor([0, 1e-1]), requires_grad=False)

    def forward(self, x):
        return self.head(x) * self.out_ubound + self.out_lbound


class RefModel:
    def __init__(self, state_dict_np):
        self.model = Net()
        self.model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in state_dict_np.items()})
        self.model.eval()
        self.model = torch.jit.script(self.model)      # model_vv.py:130

    def inference(self, batch):                        # model_vv.py:210-217
        b = torch.as_tensor(np.asarray(batch), dtype=torch.float)
        with torch.no_grad():
            out = self.model(b).cpu().split(1, dim=1)
        return [o.numpy() for o in out]
