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
        self.memory_size, self.memory_growth_rate, self.n_trains = memory_size, memory_growth_rate, 0   # ValueSim.py:21-37
        self._weights, self._rows, self.model = np.asarray(weights, np.float32), None, None
        if online and not self.benchmark:                         # ValueSim.py:21-37: the replay memory lives on the device (k_gc fills it)
            self._eng.replay_enable(min_visits=min_visits_to_store, capacity=memory_size)
            print('online: freed observations are stored on the device as in ValueSim.store_nodes (ValueSim.py:122-159); train_nodes() '
                  'trains on them (Model_VV.train_data on the device trainer)', **perr)

    def train_nodes(self, dump_data=True, path='./data/dump', **train_kwargs):    # ValueSim.py:161-185
        """Drain the device replay memory (rows k_gc stored: ValueSim.store_nodes, ValueSim.py:122-159) into ValueSim.memory; when
        memory_index >= min(n_trains * memory_growth_rate, memory_size) dump it (np.savez('./data/dump', ...), ValueSim.py:176-177), run
        Model_VV.train_data on the device trainer (ValueSim.py:180: iters_per_val=100, batch_size=1024, max_iters=50000) and hot-swap the
        search network.  Returns the memory arrays used for training, or None while still collecting (ValueSim.py:170-172)."""
        import torch
        from .. import replay
        if not (self.online and not self.benchmark):
            return None
        print('Training...', **perr)
        cap = self.memory_size
        dev = torch.device("cuda", int(self._eng.cfg.device))      # the engine's device, not torch's current one
        buf = torch.empty((cap, replay.SAMPLE_BYTES), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize(dev)                                # the engine copies on its own stream: nothing of torch's may be pending on buf
        n = self._eng.replay_drain_into(buf.data_ptr(), cap)
        rows = buf[:n].cpu().numpy()
        self._rows = rows if self._rows is None else np.concatenate([self._rows, rows])[:cap]
        d_size = len(self._rows)
        m_size = min(self.n_trains * self.memory_growth_rate, self.memory_size)
        if d_size >= m_size and d_size > 0:
            print('Enough training data ({} >= {}), proceed to training.'.format(d_size, m_size), **perr)
        else:
            print('Not enough training data ({} < {}), collecting more data.'.format(d_size, m_size), **perr)
            return None
        if dump_data:
            replay.dump(path, self._rows)
        memory = replay.rows_to_memory(self._rows)
        self.n_trains += 1
        if self.model is None:
            from ..model.model_vv import Model_VV
            self.model = Model_VV(device=int(self._eng.cfg.device), eval_kind="net")
            self.model.weights = self._weights
            self.model._eng.load_weights(self._weights)
        kw = dict(iters_per_val=100, batch_size=1024, max_iters=50000)                    # ValueSim.py:180
        kw.update(train_kwargs)
        self.model.train_data(memory, **kw)
        self.model.training(False)
        self._weights = self.model.weights
        self._eng.load_weights(self._weights)                      # the search now evaluates with the trained network
        self._rows = None                                          # memory_index = 0 (ValueSim.py:183)
        print('Training complete.', **perr)
        return memory

    def evaluate_state(self, state):                              # ValueSim.py:46-50
        v, var = self._eng.valuenet(state[None])
        return v[0], var[0]
