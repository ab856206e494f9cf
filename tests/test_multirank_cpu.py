"""CPU, world_size 2 over gloo: the host-side sharding of the round (client i -> rank i mod G, globally normalised
FedAvg weights, one all-reduce of the partial sums) reproduces the single-process aggregate.  The kernels themselves
need a GPU; here each rank forms its partial sum with the oracle."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import flb200  # noqa: F401
from flb200.simulation import global_fedavg_weights, shard_clients
from oracle import fedavg as OF


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, K, P, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    theta = (rng.standard_normal((K, P)) * 0.01).astype(np.float32)
    ns = rng.integers(100, 1000, K).tolist()
    ids = shard_clients(K, rank, world)
    w = global_fedavg_weights(ns, ids)
    partial = torch.from_numpy(OF.weighted_average_flat(theta[ids], w))
    dist.all_reduce(partial, op=dist.ReduceOp.SUM)
    full = OF.weighted_average_flat(theta, OF.sample_weights(ns))
    q.put((rank, float(np.abs(partial.numpy() - full).max()), float(np.abs(full).max()), ids, sum(w)))
    dist.destroy_process_group()


def test_two_rank_sharded_fedavg_equals_single_process():
    K, P, world = 7, 5000, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, K, P, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seen = []
    wsum = 0.0
    for rank, err, mx, ids, ws in res:
        assert err <= 1e-6 * mx            # different summation order only
        seen += ids
        wsum += ws
    assert sorted(seen) == list(range(K)) and abs(wsum - 1.0) < 1e-12


def test_shard_helpers():
    assert shard_clients(10, 1, 4) == [1, 5, 9]
    assert shard_clients(3, 3, 4) == []
    w = global_fedavg_weights([100, 300, 600], [0, 2])
    assert w == [0.1, 0.6]
