"""torchrun --nproc-per-node N scripts/p2p_bench.py : the fused peer-memory FedAvg + reduce kernel against the FedAvg kernel +
NCCL all_reduce, device-timed (CUDA events, max over ranks), for the two models' parameter counts and the sweep sizes."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flb200  # noqa: E402,F401
from flb200 import ops  # noqa: E402
from flb200.p2p import PeerFedAvg  # noqa: E402


def timed(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(40_000_000)          # ~20 ms: the host enqueues all iterations behind it, so the events see device time only
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t) * 1e3          # us


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    rows = []
    for name, P, K in (("simple_cnn", 421642, 10), ("cifar10_cnn", 1470890, 13), ("P=10M", 10_000_000, 10), ("P=50M", 50_000_000, 4)):
        ld = (P + 31) // 32 * 32
        theta = torch.randn((K, ld), device=dev)
        w = torch.full((K,), 1.0 / (K * world), device=dev)
        red = PeerFedAvg(ld, dev, rank, world, dist.group.WORLD)
        out = torch.zeros(ld, device=dev)

        def nccl():
            ops.fedavg_weighted_sum(theta, w, P=P, out=out[:P])
            dist.all_reduce(out[:P])

        def nccl_only():
            dist.all_reduce(out[:P])

        def fused():
            red.reduce(theta, w, P)

        def fedavg_only():
            ops.fedavg_weighted_sum(theta, w, P=P, out=out[:P])

        r = {"case": name, "P": P, "K": K, "world": world, "fedavg_us": timed(fedavg_only), "nccl_allreduce_us": timed(nccl_only),
             "fedavg_plus_nccl_us": timed(nccl), "fused_p2p_us": timed(fused)}
        nccl()
        a = out[:P].clone()
        b = red.reduce(theta, w, P).clone()
        r["max_abs_diff_vs_nccl"] = float((a - b).abs().max())
        rows.append(r)
        red.close()
        del red
    if rank == 0:
        for r in rows:
            print(json.dumps(r))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
