#!/usr/bin/env python
"""BASELINE.json configs[4]: FedAvg aggregation-only sweep, K client rows x P parameters (mirrors the reference's
un-run harness benchmark_aggregation_performance, src/aggregation/fedavg.py:487-548): achieved HBM GB/s of the FedAvg
kernel (fp32 and uint8-quantised inputs) and of the update-level DP kernels against the measured copy bandwidth.

    python scripts/fedavg_sweep.py [--out gpurun_out/fedavg_sweep.json]

Algorithmic bytes: FedAvg fp32 = 4*P*(K+1); q8 = P*K + 4*P; DP = 12 B/param/client (+4 for the norm pass,
which re-reads the local row).  Inputs are larger than L2 except in the smallest cells, which flush L2 between launches.
Under torchrun (N ranks) the K rows are sharded over the ranks and one NCCL all-reduce follows (time = max over ranks)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import flb200  # noqa: E402,F401
from flb200 import ops  # noqa: E402


def peak_gbs():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def timed(fn, flush, reps=5, warm=3):
    for _ in range(warm):
        fn()
    tot = 0.0
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/fedavg_sweep.json")
    ap.add_argument("--max-gb", type=float, default=120.0)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    peak, src = peak_gbs()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for P in (1_000_000, 10_000_000, 100_000_000):
        for K in (10, 100, 1000):
            Kl = len(range(rank, K, world))
            gb = Kl * P * 4 / 1e9
            if gb > args.max_gb:
                rows.append({"K": K, "P": P, "skipped": f"{gb:.0f} GB of client rows per GPU > {args.max_gb:.0f} GB"})
                continue
            ld = (P + 31) // 32 * 32
            theta = torch.empty((Kl, ld), dtype=torch.float32, device=dev).normal_(0, 0.01)     # theta ~ N(0, 0.01), fedavg.py:505-523
            g = torch.Generator().manual_seed(7)
            ns = torch.randint(100, 1000, (K,), generator=g).tolist()
            w = [ns[i] / sum(ns) for i in range(rank, K, world)]
            wt = ops.as_weight_tensor(w, dev)
            out = torch.empty(P, dtype=torch.float32, device=dev)

            def run():
                ops.fedavg_weighted_sum(theta, wt, P=P, out=out)
                if world > 1:
                    torch.distributed.all_reduce(out)
            small = Kl * P * 4 < (512 << 20)
            ms = timed(run, flush if small else None)
            if world > 1:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
                ms = float(t)
            alg = 4.0 * P * (K + world)               # every row once + one output per rank
            row = {"K": K, "P": P, "n_gpus": world, "fedavg_ms": ms, "fedavg_GBs": alg / ms / 1e6,
                   "fedavg_frac_of_peak": alg / ms / 1e6 / (peak * world)}
            if world == 1 and Kl * P <= 30e9:
                # uint8-quantised rows (QuantizationCompressor, compression.py:203-244), dequant fused into the FedAvg read
                seg = torch.tensor([0, P], dtype=torch.int64, device=dev)
                q, scale, zp = ops.q8_quantize(theta, seg, P=P)
                ms_q = timed(lambda: ops.fedavg_weighted_sum_q8(q, scale, zp, seg, wt, P), flush if small else None)
                row.update(q8_ms=ms_q, q8_GBs=(P * K + 4.0 * P) / ms_q / 1e6, q8_frac_of_peak=(P * K + 4.0 * P) / ms_q / 1e6 / peak)
                del q
            if world == 1 and Kl * P * 8 / 1e9 <= args.max_gb:
                # update-level DP: norm pass + clip/noise pass (privacy.py:107-144, 183-254)
                glob = torch.zeros(ld, dtype=torch.float32, device=dev)
                up = torch.empty_like(theta)
                ms_dp = timed(lambda: ops.dp_clip_noise(theta, glob, 1.0, 4.8448, seed=1, P=P, out=up), flush if small else None)
                row.update(dp_ms=ms_dp, dp_GBs=16.0 * P * K / ms_dp / 1e6, dp_frac_of_peak=16.0 * P * K / ms_dp / 1e6 / peak)
                del up
            rows.append(row)
            del theta, out
            torch.cuda.empty_cache()
    if rank == 0:
        res = {"peak_GBs": peak, "peak_source": src, "n_gpus": world, "rows": rows}
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        json.dump(res, open(args.out, "w"), indent=1)
        for r in rows:
            print(json.dumps(r))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
