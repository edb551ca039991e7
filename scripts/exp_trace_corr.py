"""Is a game's trace length predictable from its previous one?  (Would a separate lane for the deepest games shorten the main lane's longest walk?)
  python scripts/exp_trace_corr.py 16384 16384 500 9"""
import sys
import numpy as np
sys.path.insert(0, '.')
from tetris_mcts_b200 import pyTetris as PT, _lib as L
from tetris_mcts_b200.engine import BatchedEngine
from tetris_mcts_b200.model.model_vv import init_weights
G, M, sims, moves = (int(x) for x in sys.argv[1:5])
eng = BatchedEngine(G, max_nodes=M, mode='lp', eval_kind='net_tc', weights=init_weights(0), overflow_reset=True)
eng.set_games(PT.new_games(G, (1, 0, 0), np.arange(123, 123 + G, dtype=np.uint32)))
eng.set_gc_headroom(M * 5 // 32)
L.lib().b200_debug_trace_lens.argtypes = [L.P, L.P]
def lens():
    tl = np.zeros(G, np.int32); L.check(L.lib().b200_debug_trace_lens(eng.h, L.ptr(tl))); return tl
for mv in range(moves - 1):
    eng.play_move(sims, True, False)
# last move: sample the trace lengths inside the move
eng.run_sims(sims - 60)
prev = lens()
hist = []
for s in range(6):
    eng.run_sims(10)
    hist.append(lens())
cur = np.stack(hist)                                  # [6, G]
print('max trace since run_sims began', eng.counters()['max_trace_len'])
print('corr(previous sample, sample +10 sims) = %.3f   corr(+10, +60) = %.3f' % (np.corrcoef(prev, cur[0])[0, 1], np.corrcoef(cur[0], cur[5])[0, 1]))
for frac in (0.005, 0.01, 0.02, 0.05):
    k = int(G * frac)
    order = np.argsort(-prev)
    deep, main = order[:k], order[k:]
    print('deep lane = top %.1f %% by the previous sample (%d games, length >= %d): later samples: max over ALL games %s | max over the main lane %s | p99.9 main %s' % (
        100 * frac, k, prev[deep].min(), cur.max(1).tolist(), cur[:, main].max(1).tolist(), [int(np.percentile(c[main], 99.9)) for c in cur]))
