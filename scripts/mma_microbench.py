#!/usr/bin/env python
"""Sweep flb_mma_microbench (csrc/mma_microbench.cu): SM cycles per tcgen05.mma kind::tf32 128 x N x 8 instruction for the
operand placements the training kernels use.  Prints one JSON line per case; `floor` = 128 * N / 256 cycles (the tcgen05
dispatch floor of B300_MICROARCH.md), `smem` = operand bytes / 128 B per cycle.

    python scripts/mma_microbench.py > gpurun_out/mma_microbench.jsonl
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import flb200  # noqa: E402,F401
from flb200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
L.ensure_device(dev)
out = torch.zeros(1, dtype=torch.int64, device=dev)
reps = 512
for n in (32, 64, 128, 256):
    for a_shift, a_mn, b_mn, rotate, what in ((0, 0, 0, 1, "A K-major aligned, B K-major (conv fwd, streamed)"),
                                              (1, 0, 0, 1, "A window shifted by 1 row"),
                                              (1, 0, 0, 9, "A nine shifted windows (halo conv fwd)"),
                                              (0, 0, 0, 9, "A nine aligned windows"),
                                              (0, 0, 1, 1, "A K-major aligned, B MN-major (conv dgrad, streamed)"),
                                              (1, 0, 1, 9, "A nine shifted windows, B MN-major (halo conv dgrad)"),
                                              (0, 1, 1, 1, "A, B MN-major (wgrad)"),
                                              (1, 0, 0, 8, "UNROLLED issue loop, M = 128, eight shifted windows x 4 k-steps, B K-major"),
                                              (1, 0, 1, 8, "UNROLLED issue loop, M = 128, B MN-major"),
                                              (1, 0, 0, 7, "UNROLLED issue loop, M = 64, B K-major")):
        best = None
        for _ in range(5):
            L.call("flb_mma_microbench", n, a_shift, a_mn, b_mn, rotate, reps, L.ptr(out), L.stream_ptr(dev))
            torch.cuda.synchronize()
            c = int(out.item())
            best = c if best is None else min(best, c)
        print(json.dumps({"n": n, "case": what, "a_shift": a_shift, "a_mn": a_mn, "b_mn": b_mn, "rotate": rotate, "reps": reps,
                          "cycles_per_mma": round(best / reps, 2), "floor": 128 * n / 256, "smem_cycles": (128 * 32 + n * 32) / 128}))
