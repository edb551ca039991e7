"""Phase split of the distributional configuration (BASELINE configs[4]), development aid:  python scripts/exp_dist_phases.py 2048 32768 1500 3"""
import os, sys, time
import numpy as np
sys.path.insert(0, '.')
from tetris_mcts_b200 import pyTetris as PT
from tetris_mcts_b200.engine import BatchedEngine
from tetris_mcts_b200.agents.DistValueSimOnline import init_dist_weights
G, M, sims, moves = (int(x) for x in sys.argv[1:5])
eng = BatchedEngine(G, max_nodes=M, mode='dist', eval_kind=os.environ.get("DIST_EVAL", "net_tc"), dist_weights=init_dist_weights(0, 50), overflow_reset=True)
eng.set_games(PT.new_games(G, (1, 0, 0), np.arange(123, 123 + G, dtype=np.uint32)))
eng.set_gc_headroom(M * 5 // 32)
for timing in (False, True):
    eng.set_timing(timing)
    p0 = {k: v[0] for k, v in eng.phase_ms().items()}
    for mv in range(moves):
        eng.sync(); t = time.time()
        eng.play_move(sims, True, False)
        eng.sync(); dt = time.time() - t
        c = eng.counters()
        print('timing', timing, 'move', mv, '%.1f ms  %.3g sims/s' % (dt * 1e3, G * sims / dt), 'max_trace', c['max_trace_len'], flush=True)
    if timing:
        p1 = {k: v[0] for k, v in eng.phase_ms().items()}
        print({k: round((p1[k] - p0[k]) / moves, 1) for k in p1}, 'ms per move;  per step us:', {k: round((p1[k] - p0[k]) / moves / sims * 1e3, 1) for k in p1})
