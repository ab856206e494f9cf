"""One q8 FedAvg launch at K = 100, P = 10 M for an `ncu --set full` capture (which pipe bounds the kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flb200  # noqa
from flb200 import ops
dev = torch.device("cuda", 0)
K, P = 100, 10_000_000
ld = (P + 31) // 32 * 32
theta = torch.empty((K, ld), device=dev).normal_(0, 0.01)
seg = torch.tensor([0, P], dtype=torch.int64, device=dev)
q, scale, zp = ops.q8_quantize(theta, seg, P=P)
w = ops.as_weight_tensor([1.0 / K] * K, dev)
for _ in range(3):
    out = ops.fedavg_weighted_sum_q8(q, scale, zp, seg, w, P)
torch.cuda.synchronize()
