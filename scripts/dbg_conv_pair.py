"""Development aid: per-board error of the tensor-core value net against the CUDA-core path (which board of a CTA's pair goes wrong)."""
import sys
import numpy as np
sys.path.insert(0, '.')
from tetris_mcts_b200.engine import BatchedEngine
from tetris_mcts_b200.model.model_vv import init_weights
w = init_weights(0)
rng = np.random.default_rng(0)
def boards(n):
    b = (rng.random((n, 20, 10)) < 0.4).astype(np.int8)
    for i in range(n):
        b[i, :rng.integers(2, 12)] = 0
        r, c = rng.integers(0, 3), rng.integers(0, 8)
        b[i, r:r + 2, c:c + 2] = -1
    return b
et = BatchedEngine(1, max_nodes=64, eval_kind='net_tc', weights=w)
es = BatchedEngine(1, max_nodes=64, eval_kind='net', weights=w)
for n in (1, 2, 3, 4, 5, 8, 16, 600, 1200):
    s = boards(n)
    v, var = et.valuenet(s); v2, var2 = es.valuenet(s)
    d = np.abs(v - v2)
    print('n', n, 'max diff', d.max(), 'bad boards', np.nonzero(d > 1e-4)[0][:24].tolist(), 'n_bad', int((d > 1e-4).sum()), flush=True)
