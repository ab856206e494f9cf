#!/usr/bin/env python
"""bench.py -- one "step" = one full FedAvg round of the hot path on synthetic MNIST-shaped data:
every resident client trains 1 local epoch of the SimpleCNN (batch 32, Adam 1e-3) from the global model, the
update-level DP clip + Gaussian noise is applied to every client's delta, and the updates are FedAvg-aggregated
(NCCL all-reduce across ranks when N > 1).  Workload at N = 1 = BASELINE.json configs[1] (10 clients, DP, 1 B200);
weak scaling: 10 clients per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3                  (this framework)
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1  (CPU restatement of the reference path)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

MODEL = "simple_cnn"
CLIENTS_PER_GPU = 10
METRIC = "DP-SGD client samples/s (one FedAvg round: local epoch + update-level DP + aggregation)"
# workloads (BASELINE.json configs): the default is configs[1]; the others are extra measurements for profiles/
WORKLOADS = {
    "mnist_dp": dict(model="simple_cnn", clients_per_gpu=10, total_clients=None, compression=None, scaling="weak",
                     desc="configs[1]: SimpleCNN MNIST-shaped 28x28, 10 clients/GPU x 1 local epoch (batch 32, Adam 1e-3, dropout 0.25), "
                          "update-level DP (eps=1, delta=1e-5, C=1), FedAvg"),
    "mnist_dp50": dict(model="simple_cnn", clients_per_gpu=None, total_clients=50, compression=None, scaling="strong",
                       desc="configs[2]: SimpleCNN MNIST-shaped, 50 clients sharded over the GPUs x 1 local epoch (batch 32, Adam 1e-3, "
                            "dropout 0.25), update-level DP, FedAvg"),
    "cifar_dp_q8": dict(model="cifar10_cnn", clients_per_gpu=None, total_clients=100, compression="q8", scaling="strong",
                        desc="configs[3]: CIFAR10CNN 32x32x3, 100 clients sharded over the GPUs x 1 local epoch (batch 32, Adam 1e-3, "
                             "dropout 0.3), update-level DP, uint8-quantised updates, FedAvg"),
}
# fwd + bwd algorithmic FLOPs per sample of the GEMM-shaped kernels (2 * MAC)
GEMM_FLOPS = {
    "simple_cnn": {"conv2_fwd": 2 * 196 * 64 * 288, "conv2_dgrad": 2 * 196 * 32 * 576, "conv2_wgrad": 2 * 196 * 64 * 288, "conv2_wgrad_norm": 2 * 196 * 64 * 288,
                   "fc1_fwd": 2 * 3136 * 128, "fc1_dgrad": 2 * 3136 * 128, "fc1_wgrad": 2 * 3136 * 128},
    "cifar10_cnn": {f"conv{i}_{kind}": 2 * hw * ci * 9 * co for i, (hw, ci, co) in
                    enumerate([(1024, 3, 32), (1024, 32, 32), (256, 32, 64), (256, 64, 64), (64, 64, 128), (64, 128, 128)], start=1)
                    for kind in ("fwd", "dgrad", "wgrad") if not (i == 1 and kind == "dgrad")},
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """SM clocks / throttle reasons sampled DURING the timed regions: NVML when importable (sub-millisecond per sample),
    else the nvidia-smi query of the profiling recipe."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False      # rows: [sm_mhz, sm_max_mhz, hw, hw_thermal, sw_thermal, sw_power]
        self.nvml_error = None

    def _nvml_loop(self) -> bool:
        try:
            import pynvml as N
            N.nvmlInit()
            # CUDA_VISIBLE_DEVICES remapping: torch index -> NVML handle through the PCI bus id
            h = None
            try:
                bus = int(torch.cuda.get_device_properties(self.index).pci_bus_id)
                for i in range(N.nvmlDeviceGetCount()):
                    hh = N.nvmlDeviceGetHandleByIndex(i)
                    if int(N.nvmlDeviceGetPciInfo(hh).bus) == bus:
                        h = hh
                        break
            except Exception:
                h = None
            if h is None:
                h = N.nvmlDeviceGetHandleByIndex(self.index)
            mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
            bits = [N.nvmlClocksEventReasonHwSlowdown, N.nvmlClocksEventReasonHwThermalSlowdown,
                    N.nvmlClocksEventReasonSwThermalSlowdown, N.nvmlClocksEventReasonSwPowerCap]
            while not self.stop_flag:
                r = N.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.rows.append([N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM), mx] + [bool(r & b) for b in bits])
                time.sleep(0.005)
            return True
        except Exception as e:           # fall back to nvidia-smi; keep the reason for the record
            self.nvml_error = f"{type(e).__name__}: {e}"
            return False

    def run(self):
        if self._nvml_loop():
            return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                c = [v.strip() for v in out.split(",")]
                if len(c) >= 6 and c[0].isdigit():
                    self.rows.append([int(c[0]), int(c[1]) if c[1].isdigit() else None] + [v.lower().startswith("active") for v in c[2:6]])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "nvml_error": self.nvml_error}
        sm = sorted(r[0] for r in self.rows)
        reasons = [n for i, n in enumerate(self.NAMES) if any(r[2 + i] for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.rows[0][1], "reasons": reasons, "samples": len(self.rows)}


def cpu_baseline_run(n_clients: int, threads: int, rounds: int = 1, model: str = "simple_cnn"):
    """The reference path restated on the CPU (oracle/round.py: LocalTrainer loop + update-level DP + FedAvg), on a
    bounded sample of the same workload.  Returns (samples/s, sample description)."""
    from oracle import models as OM
    from oracle import round as OR
    MODEL = model
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    w0 = OM.init_weights(MODEL, 0)
    data = [OR.synthetic_client_data(MODEL, c) for c in range(n_clients)]
    OR.federated_round(MODEL, w0, 1, dp=True, data=data[:1], dropout_rate=0.0)          # untimed warm-up client (oneDNN init)
    t0 = time.perf_counter()
    n = 0
    for _ in range(rounds):
        _, info = OR.federated_round(MODEL, w0, n_clients, dp=True, data=data, dropout_rate=0.0)
        n += sum(info["num_samples"])
    dt = time.perf_counter() - t0
    return n / dt, f"{rounds} round(s) x {n_clients} clients x 1 local epoch (batch 32, Adam) + update-level DP + FedAvg, {n} samples in {dt:.1f} s"


def run_reference(args, out):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals = []
    sample = ""
    for i in range(args.warmup + args.steps):
        v, sample = cpu_baseline_run(CLIENTS_PER_GPU, threads, 2)
        if i >= args.warmup:
            vals.append(v)
    v = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS["mnist_dp"]["desc"], "clients": CLIENTS_PER_GPU, "dp_mode": "update", "precision": "fp32",
                       "note": "CPU restatement of the reference path (oracle/round.py) on all host threads; each step = 2 rounds "
                               "of the 10-client workload (bounded sample); dropout off (identity at p=0 costs the same)"},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=out, flush=True)


def _claim_stdout():
    """Keep stdout to the ONE JSON line: libraries (NCCL's version banner, ...) that write to fd 1 are sent to stderr;
    returns a file object on the original stdout."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("FLB_PRECISION", "tf32"), choices=["fp32", "tf32"])
    ap.add_argument("--dp-mode", default="update", choices=["update", "per_sample", "none"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="mnist_dp", choices=list(WORKLOADS))
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    MODEL = wl["model"]
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args, out)

    import flb200  # noqa: F401
    from flb200.models_pytorch import ModelFactory
    from flb200.simulation import FederatedRoundEngine, synthetic_client_data, synthetic_num_samples

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pg = None
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    n_clients = wl["total_clients"] or wl["clients_per_gpu"] * world

    eng = FederatedRoundEngine(MODEL, n_clients, dev, rank=rank, world_size=world, process_group=pg, batch_size=32,
                               local_epochs=1, learning_rate=1e-3, optimizer_type="adam", dp_mode=args.dp_mode,
                               epsilon=1.0, delta=1e-5, max_grad_norm=1.0, dropout_rate=None, precision=args.precision,
                               compression=wl["compression"])
    torch.manual_seed(0)
    w0 = ModelFactory.create_model(MODEL).get_model_weights()
    eng.set_global_weights(w0)
    host = [synthetic_client_data(MODEL, i) for i in eng.client_ids]
    sizes_all = [synthetic_num_samples(MODEL, i) for i in range(n_clients)]
    # the round's inputs as the caller holds them: this rank's clients back to back in pinned host memory
    x_host = torch.cat([h[0].reshape(h[0].shape[0], -1) for h in host]).pin_memory()
    y_host = torch.cat([h[1] for h in host]).to(torch.int32).pin_memory()
    gw_host = torch.empty(eng.layout.P, dtype=torch.float32).pin_memory()
    eng.load_packed(x_host, y_host, sizes_all)
    samples_round_local = eng.samples_per_round()
    samples_round = sum(sizes_all)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    def timed_rounds(n, e2e=False):
        """device time (CUDA events on the launch stream) summed over n rounds, L2 flushed before each"""
        total = 0.0
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if e2e:
                # H2D from pinned memory, double-buffered: this round consumes the upload issued during the previous
                # one and issues the next one on the copy stream (one 16 MB upload inside every timed step)
                # (the upload is issued BEFORE the round is enqueued: issued after it, the copy engine only gets to it when
                # the captured epoch has drained -- measured +0.15 ms per round, scripts/dbg_e2e.py)
                eng.use_prefetched()
                eng.prefetch_packed(x_host, y_host)                                         # H2D of the next round's samples
                out = eng.run_round(read_metrics=True, model_out=gw_host)                   # D2H: metrics + aggregated model, one sync
            else:
                eng.run_round(read_metrics=False)
            e1.record()
            e1.synchronize()
            total += e0.elapsed_time(e1)
        return total

    for _ in range(args.warmup):
        eng.run_round(read_metrics=False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms = timed_rounds(args.steps)
    barrier()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t.item())
    value = samples_round * args.steps / (ms / 1e3)

    # end to end through the public API: host buffers in, aggregated model + metrics out, every step
    eng.prefetch_packed(x_host, y_host)
    timed_rounds(3, e2e=True)           # both sample buffers (and their captured graphs) warm
    barrier()
    ms_e2e = timed_rounds(args.steps, e2e=True)
    barrier()
    sampler.stop_flag = True          # clocks sampled across both timed regions (device-resident and end-to-end)
    sampler.join(timeout=10)
    t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_e2e = float(t.item())
    h2d = eng.trainer.h2d_bytes
    d2h = eng.layout.P * 4 + 4 * 8 * len(eng.client_ids)

    # per-kernel breakdown of one step, CUDA events on the launch stream (eager, outside the timed region)
    tr = eng.trainer
    tr.set_global_row(eng.global_row)
    tr._fill_args(eng.lr, eng.optimizer_type, train=True)
    import ctypes as C
    from flb200 import _lib as L
    L.call("flb_train_begin_epoch", C.byref(tr.args), L.stream_ptr(dev))
    tr.profile_step()
    samples = {}
    reps = 7
    for _ in range(reps):
        for k, v in tr.profile_step().items():
            samples.setdefault(k, []).append(v)
    acc = {k: sorted(v)[len(v) // 2] for k, v in samples.items()}          # median: one preempted launch must not pick the kernel
    step_ms = sum(acc.values())
    top = max(acc, key=acc.get)
    pk = peaks()
    K_local, B = len(eng.client_ids), 32
    P = eng.layout.P
    # ALGORITHMIC work per launch (all resident clients, full batches): FLOPs of the GEMM-shaped kernels,
    # bytes of the memory-bound ones (every operand once)
    flops = GEMM_FLOPS[MODEL]
    hbm_bytes = {"optimizer": 28.0 * P * K_local}                       # Adam: read g, m, v, w; write m, v, w
    if MODEL == "simple_cnn":
        hbm_bytes.update({"conv1_fwd_pool": (3136 + 25088 + 6272.0) * B * K_local,       # x in; pooled NHWC + argmax out
                          "unpool2": (3136 * 9 + 256 * 64 * 4.0) * B * K_local,           # da2, a2, idx2 in; dz2 grid out
                          "conv1_wgrad": (3136 + 25088 * 2 + 6272.0) * B * K_local})      # x, a1p, da1p, idx1 in
    tf32_peak = pk["bf16_tflops_sustained"] / 2.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")           # dram bytes per launch from `ncu --set full`
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(MODEL, {}).get(str(K_local), {}).get(top)
    if top in flops:
        ach = flops[top] * B * K_local / (acc[top] * 1e-3) / 1e12
        roof = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach / tf32_peak,
                "traffic": traffic, "peak_source": "0.5 x sustained bf16 of " + pk["source"] + " (TF32 = half the bf16 rate)"}
    elif top in hbm_bytes:
        ach = hbm_bytes[top] / (acc[top] * 1e-3) / 1e9
        roof = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                "traffic": traffic, "peak_source": pk["source"], "algorithmic_bytes_per_launch": hbm_bytes[top]}
    else:
        roof = {"kernel": top, "bound": "hbm", "achieved": None, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": None, "traffic": traffic}
    roof["share_of_step"] = acc[top] / step_ms
    roof["how"] = ("CUDA events after every kernel of one training step, launched eagerly on the launch stream right after the "
                   "timed region (the timed rounds replay a CUDA graph, which cannot be instrumented per kernel); median of 7 steps")
    roof["step_breakdown_ms"] = {k: round(v, 5) for k, v in acc.items()}
    # the other kernels against their own rooflines, for the record
    others = {}
    for name, ms_k in acc.items():
        if name in flops:
            others[name] = {"TFLOP/s": round(flops[name] * B * K_local / (ms_k * 1e-3) / 1e12, 2), "frac_tensor": round(flops[name] * B * K_local / (ms_k * 1e-3) / 1e12 / tf32_peak, 4)}
        elif name in hbm_bytes:
            others[name] = {"GB/s": round(hbm_bytes[name] / (ms_k * 1e-3) / 1e9, 1), "frac_hbm": round(hbm_bytes[name] / (ms_k * 1e-3) / 1e9 / pk["hbm_gbs"], 4)}
    roof["per_kernel"] = others

    # our kernels per round: the epoch's launch sequence + update-level DP (norm, clip+noise) + FedAvg (vector body, scalar tail)
    launches_round = tr.launches_per_epoch() + (2 if args.dp_mode == "update" else 0) + 2
    line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": "tf32" if args.precision == "tf32" else "f32", "data": "synthetic",
            "config": {"workload": wl["desc"] + (" + NCCL all-reduce" if world > 1 else ""),
                       "clients": n_clients, "samples_per_round": samples_round, "dp_mode": args.dp_mode, "precision": args.precision,
                       "l2": "flushed (256 MB write) before every timed round", "round_ms": ms / args.steps},
            "e2e": {"value": samples_round * args.steps / (ms_e2e / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_round * args.steps,
            "roofline": roof, "clocks": sampler.summary()}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        # a bounded sample of the same workload, ~10 s of CPU work: 24 rounds of the 10-client SimpleCNN job / 1 round of 8
        # CIFAR10CNN clients
        v, sample = (cpu_baseline_run(CLIENTS_PER_GPU, threads, 24) if MODEL == "simple_cnn"
                     else cpu_baseline_run(8, threads, 1, MODEL))
        line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample}
    if rank == 0:
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
