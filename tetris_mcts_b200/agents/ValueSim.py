"""agents.ValueSim — agents/ValueSim.py:12-99 (leaf itself evaluated, gamma=0.999, max_nodes=100000)."""
from sys import stderr

from .agent import TreeAgent
from ..model.model_vv import init_weights

perr = dict(file=stderr, flush=True)


class ValueSim(TreeAgent):
    _mode = "single"

    def __init__(self, online=True, memory_size=500000, min_visits_to_store=10, gamma=0.999, memory_growth_rate=5000, weights=None, **kwargs):
        kwargs.pop("max_nodes", None)
        super().__init__(max_nodes=100000, gamma=gamma, low=1, weights=init_weights(0) if weights is None else weights, **kwargs)   # ValueSim.py:16
        self.online, self.min_visits_to_store = online, min_visits_to_store
        if online and not self.benchmark:
            print('online training (ValueSim.py:101-185) is outside the accelerated path (SURVEY 8f): samples can be drawn with '
                  'BatchedEngine.collect_samples_into, no optimiser step is run here', **perr)

    def evaluate_state(self, state):                              # ValueSim.py:46-50
        v, var = self._eng.valuenet(state[None])
        return v[0], var[0]
