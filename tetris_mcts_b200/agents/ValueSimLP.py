"""agents.ValueSimLP — agents/ValueSimLP.py:7-70 (leaf-parallel: the <= 7 unique children of the leaf are evaluated in
one batch and averaged; min_visits_to_store=25)."""
from .ValueSim import ValueSim


class ValueSimLP(ValueSim):
    _mode = "lp"

    def __init__(self, **kwargs):
        super().__init__(min_visits_to_store=25, **kwargs)        # ValueSimLP.py:11
