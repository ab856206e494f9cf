"""Oracle (test infrastructure): per-sample DP-SGD step.  PARITY UNPINNED BY THE REFERENCE.

The reference never computes per-sample gradients (``opacus`` is listed in
``requirements.txt:7`` and never imported; local training is plain minibatch descent,
``src/shared/training.py:173-212``).  BASELINE.json's north-star kernel (2) therefore has no
upstream implementation; this oracle restates the textbook DP-SGD step (Abadi et al. 2016)
using the reference's own rules where it has them:
  * clip rule   ``src/shared/privacy.py:127-138``: scale by C/||g_i|| only if ||g_i|| > C
  * sigma rule  ``src/shared/privacy.py:209``:     sigma = S*sqrt(2 ln(1.25/delta))/eps, S = C
  * hook point  between ``loss.backward()`` and ``optimizer.step()``, ``training.py:196-197``
    (the unused ``get_model_gradients`` / ``set_model_gradients`` hooks ``:362-384``)
Step:  g_i = grad of sample i's loss;  gbar = (sum_i g_i * min(1, C/||g_i||) + sigma * z) / B.
Per-sample gradients come from ``torch.func.vmap(grad(functional_call))``.

CIFAR10CNN (``src/shared/models_pytorch.py:100-165``) has six BatchNorm2d layers, whose batch statistics couple the
samples of a minibatch: "the gradient of sample i" needs a rule the reference does not give.  The rule restated here
(VERDICT r1 item 5): the forward pass is the reference's own (train mode, BATCH statistics), and in
the per-sample backward pass every layer's batch mean / biased variance are CONSTANTS -- g_i is the gradient of sample
i's loss through a network whose BatchNorm layers normalise with those fixed numbers.  Running statistics are updated
by the forward pass as upstream.  Equally unpinned.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch.func import grad, vmap

from . import models as M
from .privacy import gaussian_sigma


def _per_sample_grads_frozen_bn(w: Dict[str, torch.Tensor], x: torch.Tensor, y: torch.Tensor,
                                dropout_rate: float = 0.0, masks=None):
    """CIFAR10CNN: batch statistics recorded by one train-mode forward pass of the whole minibatch, then one
    backward pass per sample with those statistics as constants (dropout masks: the sample's rows of the batch's masks)."""
    wd = {k: v.detach() for k, v in w.items()}
    stats: dict = {}
    with torch.no_grad():
        M.cifar10_cnn_forward(wd, x, train=True, dropout_rate=dropout_rate, masks=masks, bn_record=stats)
    out = {k: [] for k in wd}
    for i in range(x.shape[0]):
        wl = {k: v.clone().requires_grad_(True) for k, v in wd.items()}
        mi = [m[i:i + 1] for m in masks] if masks is not None else None
        logits = M.cifar10_cnn_forward(wl, x[i:i + 1], train=True, dropout_rate=dropout_rate, masks=mi, bn_fixed=stats)
        loss = F.cross_entropy(logits, y[i:i + 1])
        gs = torch.autograd.grad(loss, list(wl.values()))
        for k, g in zip(wl, gs):
            out[k].append(g)
    return {k: torch.stack(v) for k, v in out.items()}


def per_sample_grads(model: str, w: Dict[str, torch.Tensor], x: torch.Tensor, y: torch.Tensor,
                     dropout_rate: float = 0.0, masks=None):
    if model == "cifar10_cnn":
        return _per_sample_grads_frozen_bn(w, x, y, dropout_rate, masks)
    assert masks is None, "injected dropout masks: cifar10_cnn restatement only"

    def loss_one(wl, xi, yi):
        logits = M.forward(model, wl, xi.unsqueeze(0), train=True, dropout_rate=0.0)
        return F.cross_entropy(logits, yi.unsqueeze(0))

    return vmap(grad(loss_one), in_dims=(None, 0, 0))({k: v.detach() for k, v in w.items()}, x, y)


def per_sample_norms(g: Dict[str, torch.Tensor]) -> torch.Tensor:
    sq = None
    for t in g.values():
        s = t.reshape(t.shape[0], -1).pow(2).sum(1)
        sq = s if sq is None else sq + s
    return sq.sqrt()


def dp_sgd_grad(model: str, w, x, y, max_norm: float, epsilon: float, delta: float,
                z: Optional[Dict[str, torch.Tensor]] = None, dropout_rate: float = 0.0, masks=None):
    """Returns (gbar dict, per-sample norms [B], sigma)."""
    g = per_sample_grads(model, w, x, y, dropout_rate, masks)
    norms = per_sample_norms(g)
    coef = torch.where(norms > max_norm, max_norm / norms, torch.ones_like(norms))
    sigma = gaussian_sigma(max_norm, epsilon, delta)
    B = x.shape[0]
    out = {}
    for k, t in g.items():
        s = (t * coef.reshape(-1, *([1] * (t.dim() - 1)))).sum(0)
        if z is not None:
            s = s + sigma * z[k]
        out[k] = s / B
    return out, norms, sigma
