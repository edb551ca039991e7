"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes exercise sharding, the fixed-block sample all-gather
and the counter reductions that bench.py uses (NCCL on the GPU box)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from tetris_mcts_b200 import distributed as D


def test_shard_range_covers_everything():
    for n in (1, 7, 16384, 65536, 1000003):
        for world in (1, 2, 3, 4, 8):
            edges = [D.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    s0 = D.shard_seeds(123, 10, 0, 2)
    s1 = D.shard_seeds(123, 10, 1, 2)
    assert np.array_equal(np.concatenate([s0, s1]), np.arange(123, 133, dtype=np.uint32))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    D.init(backend="gloo")
    cap = 16
    count = 3 + 4 * rank                      # ragged: 3 and 7
    block = torch.zeros((cap, D.SAMPLE_BYTES), dtype=torch.uint8)
    for i in range(count):
        block[i, :200] = (rank * 50 + i) % 3          # fake board bytes
        f = np.array([rank * 100 + i, 0.5 * i, 25 + i], np.float32).view(np.uint8)
        block[i, 200:212] = torch.from_numpy(f.copy())
    rows, counts = D.allgather_samples(block, count)
    states, v, var, w = D.decode_samples(rows)
    tot = D.sum_over_ranks({"sims": 1000 * (rank + 1), "games": 5})
    mx = D.max_over_ranks(10.0 + rank)
    D.barrier()
    q.put((rank, counts, rows.shape[0], v.ravel().tolist(), w.ravel().tolist(), tot, mx, states.shape))
    torch.distributed.destroy_process_group()


def test_allgather_and_reductions_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, counts, n, v, w, tot, mx, sshape in res:
        assert counts == [3, 7] and n == 10 and sshape == (10, 1, 20, 10)
        assert v == [0.0, 1.0, 2.0] + [100.0 + i for i in range(7)]
        assert w[:3] == [25.0, 26.0, 27.0]
        assert tot == {"games": 10.0, "sims": 3000.0} and mx == 11.0


def test_single_process_passthrough():
    block = torch.zeros((4, D.SAMPLE_BYTES), dtype=torch.uint8)
    rows, counts = D.allgather_samples(block, 2)
    assert rows.shape == (2, D.SAMPLE_BYTES) and counts == [2]
    assert D.max_over_ranks(3.5) == 3.5
