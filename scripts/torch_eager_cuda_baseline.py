#!/usr/bin/env python
"""The "kernel to beat" of SURVEY.md section 8(d): the same round (K clients x 1 local epoch, batch 32, Adam 1e-3 ->
update-level DP clip + Gaussian noise -> FedAvg) written the way the reference would run it on a GPU -- stock eager
PyTorch (cuDNN / cuBLAS, TF32 allowed), clients one after the other, one optimizer per client, `.item()` per step as in
src/shared/training.py:200-203.  Library code only: none of this repo's kernels are on this path, and nothing here is
imported by the product.  A measurement script: prints one JSON line.

    python scripts/torch_eager_cuda_baseline.py --model simple_cnn --clients 10 --rounds 5
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flb200.models_pytorch import INPUT_SHAPES, ModelFactory  # noqa: E402  (parameter containers only)


def forward(m, x):
    if m.model_name == "simple_cnn":
        x = m.pool(F.relu(m.conv1(x)))
        x = m.pool(F.relu(m.conv2(x)))
        x = m.dropout(F.relu(m.fc1(x.flatten(1))))
        return m.fc2(x)
    for blk in range(3):
        a, b = 2 * blk + 1, 2 * blk + 2
        x = F.relu(getattr(m, f"bn{a}")(getattr(m, f"conv{a}")(x)))
        x = F.relu(getattr(m, f"bn{b}")(getattr(m, f"conv{b}")(x)))
        x = m.dropout(m.pool(x))
    x = m.dropout(F.relu(m.fc1(x.flatten(1))))
    x = m.dropout(F.relu(m.fc2(x)))
    return m.fc3(x)


def one_round(model_name, w_global, data, dev, eps=1.0, delta=1e-5, clip=1.0):
    uploads, sizes = [], []
    for x, y in data:
        m = ModelFactory.create_model(model_name).to(dev)
        m.set_model_weights(w_global)
        m.train()
        opt = torch.optim.Adam(m.parameters(), lr=1e-3)
        tot = 0.0
        correct = 0
        for i in range(0, x.shape[0], 32):
            xb, yb = x[i:i + 32], y[i:i + 32]
            opt.zero_grad()
            out = forward(m, xb)
            loss = F.cross_entropy(out, yb)
            loss.backward()
            opt.step()
            tot += loss.item()                                         # training.py:200
            correct += (out.argmax(1) == yb).sum().item()              # training.py:203
        w = m.get_model_weights()
        d = {k: w[k] - w_global[k] for k in w}                         # federated_trainer.py:441
        n = math.sqrt(sum(float(t.norm().item()) ** 2 for t in d.values()))   # privacy.py:118-125
        coef = clip / n if n > clip else 1.0
        sigma = min(n, clip) * math.sqrt(2 * math.log(1.25 / delta)) / eps
        uploads.append({k: w_global[k] + d[k] * coef + torch.normal(0.0, sigma, d[k].shape, device=dev) for k in d})
        sizes.append(x.shape[0])
    tot_n = float(sum(sizes))
    agg = {k: torch.zeros_like(v) for k, v in w_global.items()}
    for u, n_i in zip(uploads, sizes):                                 # fedavg.py:267-289
        for k in agg:
            agg[k] += (n_i / tot_n) * u[k]
    return agg, sum(sizes)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="simple_cnn")
    ap.add_argument("--clients", type=int, default=10)
    ap.add_argument("--rounds", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    sizes = [480, 512, 544, 576] if args.model == "simple_cnn" else [416, 448, 480]
    data = []
    for c in range(args.clients):
        g = torch.Generator().manual_seed(1000 + c)
        n = sizes[c % len(sizes)]
        data.append((torch.randn((n,) + INPUT_SHAPES[args.model], generator=g).to(dev),
                     torch.randint(0, 10, (n,), generator=g).to(dev)))
    torch.manual_seed(0)
    w = {k: v.to(dev) for k, v in ModelFactory.create_model(args.model).get_model_weights().items()}
    for _ in range(args.warmup):
        w2, _ = one_round(args.model, w, data, dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    for _ in range(args.rounds):
        w2, k = one_round(args.model, w, data, dev)
        n += k
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"impl": "torch_eager_cuda", "model": args.model, "clients": args.clients, "rounds": args.rounds,
                      "ms_per_round": dt / args.rounds * 1e3, "samples_per_s": n / dt,
                      "note": "stock eager PyTorch (cuDNN/cuBLAS, TF32 allowed), clients sequential, inputs resident in HBM; "
                              "wall clock around synchronize()", "torch": torch.__version__,
                      "gpu": torch.cuda.get_device_name(0)}))


if __name__ == "__main__":
    main()
