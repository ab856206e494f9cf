"""CPU: internal consistency of the per-sample DP-SGD restatement (oracle/dpsgd.py -- PARITY UNPINNED, the reference has no
per-sample code, see its header).  These checks do not pin it to the reference; they pin it to itself and to the batch gradient."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import dpsgd as OD
from oracle import models as OM
from oracle import privacy as OPV


def _data(model, seed, n):
    g = torch.Generator().manual_seed(seed)
    return torch.randn((n,) + OM.input_shape(model), generator=g), torch.randint(0, 10, (n,), generator=g)


def test_simple_cnn_per_sample_gradients_sum_to_the_batch_gradient():
    model = "simple_cnn"
    w = OM.init_weights(model, 3)
    x, y = _data(model, 1, 6)
    g = OD.per_sample_grads(model, w, x, y)
    wl = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    loss = F.cross_entropy(OM.forward(model, wl, x, train=True), y, reduction="sum")
    ref = torch.autograd.grad(loss, list(wl.values()))
    for (k, v), r in zip(g.items(), ref):
        np.testing.assert_allclose(v.sum(0).numpy(), r.numpy(), rtol=2e-4, atol=2e-6, err_msg=k)


def test_cifar_frozen_statistics_rule():
    """CIFAR10CNN: (a) the per-sample loop equals ONE batched backward pass through the network with every BatchNorm layer
    normalising with the recorded batch statistics as constants (the rule, stated once more in a different way); (b) behind
    the last BatchNorm (fc1-3) that is the ordinary batch gradient of the summed loss; (c) in front of it, it is NOT -- the
    coupling through the batch statistics is what the rule removes."""
    model = "cifar10_cnn"
    w = OM.init_weights(model, 4)
    gen = torch.Generator().manual_seed(9)
    for i in range(1, 7):
        w[f"bn{i}.weight"] = 1.0 + 0.2 * torch.randn(w[f"bn{i}.weight"].shape, generator=gen)
        w[f"bn{i}.bias"] = 0.1 * torch.randn(w[f"bn{i}.bias"].shape, generator=gen)
    x, y = _data(model, 2, 5)
    g = OD.per_sample_grads(model, w, x, y)
    stats = {}
    with torch.no_grad():
        OM.cifar10_cnn_forward(w, x, train=True, bn_record=stats)
    wl = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    loss = F.cross_entropy(OM.cifar10_cnn_forward(wl, x, train=True, bn_fixed=stats), y, reduction="sum")
    frozen = dict(zip(wl, torch.autograd.grad(loss, list(wl.values()))))
    wl2 = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    loss2 = F.cross_entropy(OM.cifar10_cnn_forward(wl2, x, train=True), y, reduction="sum")
    batch = dict(zip(wl2, torch.autograd.grad(loss2, list(wl2.values()))))
    rel = lambda a, b: float((a - b).norm() / b.norm().clamp(min=1e-30))     # noqa: E731
    for k in w:
        assert rel(g[k].sum(0), frozen[k]) < 2e-4, k                         # (a)
    for k in ("fc1.weight", "fc2.weight", "fc3.weight", "fc3.bias"):
        assert rel(g[k].sum(0), batch[k]) < 2e-4, k                          # (b)
    assert rel(g["conv1.weight"].sum(0), batch["conv1.weight"]) > 1e-2       # (c)
    assert float(batch["conv2.bias"].abs().max()) < 1e-5 < float(g["conv2.bias"].sum(0).abs().max())   # bias in front of a BatchNorm


def test_clip_and_noise_rule():
    model = "simple_cnn"
    w = OM.init_weights(model, 5)
    x, y = _data(model, 3, 8)
    norms = OD.per_sample_norms(OD.per_sample_grads(model, w, x, y))
    C = float(norms.median())
    z = {k: torch.ones_like(v) for k, v in w.items()}
    gbar, n2, sigma = OD.dp_sgd_grad(model, w, x, y, C, 1.0, 1e-5, z=z)
    assert torch.equal(n2, norms) and abs(sigma - OPV.gaussian_sigma(C, 1.0, 1e-5)) < 1e-12
    g = OD.per_sample_grads(model, w, x, y)
    coef = torch.where(norms > C, C / norms, torch.ones_like(norms))         # privacy.py:127-138 per sample
    assert (coef < 1).any() and (coef == 1).any()
    for k in w:
        ref = ((g[k] * coef.reshape(-1, *([1] * (g[k].dim() - 1)))).sum(0) + sigma) / 8
        np.testing.assert_allclose(gbar[k].numpy(), ref.numpy(), rtol=1e-6, atol=1e-7)
    clipped = OD.per_sample_norms({k: g[k] * coef.reshape(-1, *([1] * (g[k].dim() - 1))) for k in g})
    assert float(clipped.max()) <= C * (1 + 1e-5)
