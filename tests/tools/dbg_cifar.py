"""debug aid: per-tensor gradient error of the CIFAR10CNN kernels vs the CPU oracle (fp64 oracle as the arbiter)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import flb200
from flb200.training import BatchedClientTrainer
from oracle import models as OM, training as OT
MODEL = "cifar10_cnn"
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
g = torch.Generator().manual_seed(21)
x = torch.randn((n, 3, 32, 32), generator=g); y = torch.randint(0, 10, (n,), generator=g)
w = OM.init_weights(MODEL, 11)
eng = BatchedClientTrainer(MODEL, 1, torch.device("cuda:0"), batch_size=n, dropout_rate=0.0, precision=prec)
eng.set_client_weights(0, w); eng.load_data([x], [y]); eng.forward_backward()
got = {k: v.cpu().double() for k, v in eng.layout.views(eng.G[0]).items()}
loss32, logits32, g32 = OT.loss_and_grads(MODEL, w, x, y, train=True, dropout_rate=0.0)
w64 = {k: v.double() for k, v in w.items()}
loss64, logits64, g64 = OT.loss_and_grads(MODEL, w64, x.double(), y, train=True, dropout_rate=0.0)
lg = eng.ws_array("logits", torch.float32, 10)[0, :n].cpu().double()
print("logits err ours %.2e  cpu32 %.2e" % ((lg - logits64).abs().max(), (logits32.double() - logits64).abs().max()))
for k in g64:
    r = g64[k]; nr = r.norm() + 1e-30
    print("%-14s ours %.2e   cpu-fp32 %.2e   |ref| %.2e" % (k, (got[k] - r).norm() / nr, (g32[k].double() - r).norm() / nr, nr))
# ---- intermediate check: dz2, dz1 -------------------------------------------------------------------------------------
import torch.nn.functional as F
wl = {k: v.double().requires_grad_(True) for k, v in w.items()}
xd = x.double()
z1 = F.conv2d(xd, wl["conv1.weight"], wl["conv1.bias"], padding=1); z1.retain_grad()
y1 = F.relu(F.batch_norm(z1, None, None, wl["bn1.weight"], wl["bn1.bias"], True, 0.1, 1e-5)); y1.retain_grad()
z2 = F.conv2d(y1, wl["conv2.weight"], wl["conv2.bias"], padding=1); z2.retain_grad()
y2 = F.relu(F.batch_norm(z2, None, None, wl["bn2.weight"], wl["bn2.bias"], True, 0.1, 1e-5))
p1 = F.max_pool2d(y2, 2, 2); p1.retain_grad()
def block(x, i):
    return F.relu(F.batch_norm(F.conv2d(x, wl[f"conv{i}.weight"], wl[f"conv{i}.bias"], padding=1), None, None, wl[f"bn{i}.weight"], wl[f"bn{i}.bias"], True, 0.1, 1e-5))
t = F.max_pool2d(block(block(p1, 3), 4), 2, 2)
t = F.max_pool2d(block(block(t, 5), 6), 2, 2).reshape(-1, 2048)
t = F.relu(F.linear(t, wl["fc1.weight"], wl["fc1.bias"])); t = F.relu(F.linear(t, wl["fc2.weight"], wl["fc2.bias"]))
loss = F.cross_entropy(F.linear(t, wl["fc3.weight"], wl["fc3.bias"]), y); loss.backward()
def grid(name, C, PP, Wp, H):
    t = eng.ws_array(name, torch.float32, PP * C)[0, :n].cpu().double().view(n, PP, C)
    idx = (torch.arange(H).view(H, 1) * Wp + torch.arange(H).view(1, H)).reshape(-1)
    return t[:, idx, :].view(n, H, H, C).permute(0, 3, 1, 2)
for nm, ref, (C, PP, Wp, H) in (("d32a", z2.grad, (32, 1096, 33, 32)), ("d32b", z1.grad, (32, 1096, 33, 32)), ("d16p", p1.grad, (32, 296, 17, 16)),
                                ("z2", z2.detach(), (32, 1096, 33, 32)), ("y1", y1.detach(), (32, 1096, 33, 32))):
    got_t = grid(nm, C, PP, Wp, H)
    err = (got_t - ref).abs()
    print(nm, "rel err %.2e" % ((got_t - ref).norm() / ref.norm()), "max at", np.unravel_index(int(err.argmax()), err.shape), "max err %.2e" % err.max())
    e2 = err.amax(dim=(0, 1)); print("   per-row max err (first/last rows):", e2.amax(dim=1)[:3].tolist(), e2.amax(dim=1)[-3:].tolist(), " per-col:", e2.amax(dim=0)[:3].tolist(), e2.amax(dim=0)[-3:].tolist())
got_t = grid("d32a", 32, 1096, 33, 32); ref = z2.grad
err = (got_t - ref)
bad = (err.abs() > 1e-6).nonzero()
print("entries off by > 1e-6:", bad.shape[0])
for b_, c_, h_, w_ in bad.tolist()[:6]:
    h0, w0 = h_ // 2 * 2, w_ // 2 * 2
    print((b_, c_, h_, w_), "ours", got_t[b_, c_, h0:h0 + 2, w0:w0 + 2].flatten().tolist(), "ref", ref[b_, c_, h0:h0 + 2, w0:w0 + 2].flatten().tolist())
    bn = F.batch_norm(z2.detach(), None, None, wl["bn2.weight"].detach(), wl["bn2.bias"].detach(), True, 0.1, 1e-5)
    print("    bn2 window (fp64):", bn[b_, c_, h0:h0 + 2, w0:w0 + 2].flatten().tolist())
