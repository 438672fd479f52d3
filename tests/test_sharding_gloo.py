"""Multi-GPU partitioning on CPU: deterministic LPT packing, exact cover, and
a world_size-2 gloo run in which every rank derives the same plan and the results gathered on rank
0 come back in input order."""
import os
import socket
import sys

import numpy as np
import pytest

from scape_b200 import shard


def test_lpt_partition_is_an_exact_balanced_cover():
    rng = np.random.default_rng(0)
    reads = [np.clip(np.rint(np.exp(rng.normal(np.log(600), 1.6, size=100))), 10, 200000).astype(int) for _ in range(200)]
    costs = shard.stream_costs(reads)
    for g in (1, 2, 4, 8):
        parts = shard.lpt_partition(costs, g)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(200))
        assert all(p == sorted(p) for p in parts)
        assert shard.imbalance(costs, parts) < 1.05
    assert shard.lpt_partition(costs, 8) == shard.lpt_partition(costs, 8)


def test_cost_model_is_monotone():
    c = [shard.utr_cost(n) for n in (10, 100, 1000, 10000, 100000)]
    assert all(a < b for a, b in zip(c, c[1:]))
    assert shard.utr_cost(500, 20000) > shard.utr_cost(500, 2000)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    reads = [[100 + 37 * ((s * 7 + i) % 11) for i in range(5)] for s in range(9)]
    parts = shard.lpt_partition(shard.stream_costs(reads), world)
    mine = [(s, [f"stream{s}-utr{i}" for i in range(5)]) for s in parts[rank]]      # stand-in for the GPU fit
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        merged = dict(kv for part in gathered for kv in part)
        q.put([merged[s] for s in sorted(merged)])
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_gather_keeps_input_order():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out == [[f"stream{s}-utr{i}" for i in range(5)] for s in range(9)]
