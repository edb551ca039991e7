"""agents.ValueSim — agents/ValueSim.py:12-99 (leaf itself evaluated, gamma=0.999, max_nodes=100000)."""
from sys import stderr

import numpy as np

from .agent import TreeAgent
from ..model.model_vv import init_weights, load_checkpoint_weights

perr = dict(file=stderr, flush=True)


class ValueSim(TreeAgent):
    _mode = "single"

    def __init__(self, online=True, memory_size=500000, min_visits_to_store=10, gamma=0.999, memory_growth_rate=5000, weights=None, **kwargs):
        kwargs.pop("max_nodes", None)
        if weights is None:
            # ValueSim.py:42-44: self.model = Model(); self.model.load(); self.model.training(False).  Model.load (model/model.py:163-174)
            # reads ./pytorch_model/model_checkpoint when it exists and otherwise keeps the default-initialised network.
            weights = load_checkpoint_weights()
            if weights is None:
                weights = init_weights(0)
        super().__init__(max_nodes=100000, gamma=gamma, low=1, weights=weights, **kwargs)   # ValueSim.py:16
        self.online, self.min_visits_to_store = online, min_visits_to_store
        if online and not self.benchmark:                         # ValueSim.py:21-37: the replay memory lives on the device (k_gc fills it)
            self._eng.replay_enable(min_visits=min_visits_to_store, capacity=memory_size)
            print('online: samples are stored as in ValueSim.store_nodes (ValueSim.py:122-159); train_nodes() returns / dumps them, the '
                  'optimiser step itself is outside the accelerated path (SURVEY 8f.2)', **perr)

    def train_nodes(self, dump_data=True, path='./data/dump'):    # ValueSim.py:161-185, the data half: drain the stored samples
        """Returns ValueSim.memory[:d_size] (states, values, variance, weights) and writes the reference's dump file
        (ValueSim.py:176-177).  The optimiser step itself (model.train_data) is outside the accelerated path."""
        import torch
        from .. import replay
        if not (self.online and not self.benchmark):
            return replay.rows_to_memory(np.zeros((0, replay.SAMPLE_BYTES), np.uint8))
        cap = 500000
        dev = torch.device("cuda", int(self._eng.cfg.device))      # the engine's device, not torch's current one
        buf = torch.empty((cap, replay.SAMPLE_BYTES), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize(dev)                                # the engine copies on its own stream: nothing of torch's may be pending on buf
        n = self._eng.replay_drain_into(buf.data_ptr(), cap)
        rows = buf[:n].cpu().numpy()
        if dump_data and n:
            replay.dump(path, rows)
        return replay.rows_to_memory(rows)

    def evaluate_state(self, state):                              # ValueSim.py:46-50
        v, var = self._eng.valuenet(state[None])
        return v[0], var[0]
