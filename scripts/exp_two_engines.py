"""Does running the tree phase of one group of games under the network phase of another pay?  Development experiment:
K engines of G/K games each on their own streams, their moves enqueued concurrently (b200_run_sims is asynchronous: 500 graph launches),
against one engine of G games.  Same games, same seeds, same search results; only the schedule differs.

  python scripts/exp_two_engines.py 16384 16384 500 10 1,2,4
"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from tetris_mcts_b200 import pyTetris as PT
from tetris_mcts_b200.engine import BatchedEngine
from tetris_mcts_b200.model.model_vv import init_weights

G, M, sims, moves = (int(x) for x in sys.argv[1:5])
ks = [int(x) for x in sys.argv[5].split(',')]
w = init_weights(0)
for K in ks:
    g = G // K
    engs = []
    for k in range(K):
        e = BatchedEngine(g, max_nodes=M, mode='lp', eval_kind='net_tc', weights=w, overflow_reset=True)
        e.set_games(PT.new_games(g, (1, 0, 0), np.arange(123 + k * g, 123 + (k + 1) * g, dtype=np.uint32)))
        e.set_gc_headroom(M * 5 // 32)
        engs.append(e)
    per_move = []
    for mv in range(moves):
        for e in engs:
            e.sync()
        t = time.time()
        for e in engs:
            e.run_sims(sims)                 # asynchronous
        for e in engs:                       # play_move's tail (b200_play_move): statistics -> env step -> re-root
            e.get_stats()
            e.env_step(None)
            e.update_root(True)
        for e in engs:
            e.sync()
        per_move.append(time.time() - t)
    tail = per_move[5:] if moves > 5 else per_move
    c = [e.counters() for e in engs]
    print('K=%d engines x %d games: ms/move %s | moves 5+: %.1f ms/move = %.3g sims/s | sims %d finished %d' % (
        K, g, ' '.join('%.0f' % (x * 1e3) for x in per_move), 1e3 * np.mean(tail), G * sims / np.mean(tail),
        sum(x['sims'] for x in c), sum(x['games_finished'] for x in c)), flush=True)
    for e in engs:
        e.close()
    del engs
