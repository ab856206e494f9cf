"""Oracle (test infrastructure): update-level differential privacy, restated.

Follows ``src/shared/privacy.py``:
  * ``GradientClipper.clip_gradients`` ``:107-144`` -- global L2 norm over ALL tensors
    (per-tensor fp32 ``norm()`` squared and summed in Python float ``:119-123``), scale by
    ``C / norm`` only when ``norm > C`` (``:127-131``) else clone (``:137-138``);
    returns ``(clipped, min(norm, C))`` (``:140``).
  * ``GaussianNoiseGenerator.generate_noise`` ``:183-219`` -- sigma rule ``:209``.
  * ``add_noise_to_gradients`` ``:221-254`` -- per tensor, dict order, ``grad + noise``.
  * ``DifferentialPrivacyEngine.add_noise`` ``:284-311`` -- clip, then noise with
    ``sensitivity = min(norm, C)``.
The noise itself is INJECTED (a dict of standard-normal tensors ``z``; noise = sigma * z) so
the CUDA path can be compared with an identical noise tensor, or drawn with
``torch.normal`` in dict order exactly like the reference when ``z`` is None.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch


def gaussian_sigma(sensitivity: float, epsilon: float, delta: float) -> float:
    # privacy.py:202-209
    if epsilon <= 0:
        raise ValueError("Epsilon must be positive")
    if delta <= 0 or delta >= 1:
        raise ValueError("Delta must be in (0, 1)")
    return sensitivity * math.sqrt(2 * math.log(1.25 / delta)) / epsilon


def global_norm(g: Dict[str, Optional[torch.Tensor]]) -> float:
    # privacy.py:119-123
    acc = 0.0
    for t in g.values():
        if t is not None:
            acc += t.norm().item() ** 2
    return math.sqrt(acc)


def clip(g: Dict[str, Optional[torch.Tensor]], max_norm: float) -> Tuple[Dict[str, torch.Tensor], float]:
    # privacy.py:107-144
    n = global_norm(g)
    if n > max_norm:
        coef = max_norm / n
        out = {k: (t * coef if t is not None else t) for k, t in g.items()}
    else:
        out = {k: (t.clone() if t is not None else t) for k, t in g.items()}
    return out, min(n, max_norm)


def add_noise(g: Dict[str, torch.Tensor], epsilon: float, delta: float, max_norm: float,
              z: Optional[Dict[str, torch.Tensor]] = None) -> Tuple[Dict[str, torch.Tensor], float, float]:
    """privacy.py:284-311 (budget bookkeeping excluded).  Returns (noisy, sensitivity, sigma)."""
    clipped, sens = clip(g, max_norm)
    sigma = gaussian_sigma(sens, epsilon, delta)
    out = {}
    for k, t in clipped.items():
        if t is None:
            out[k] = t
        elif z is None:
            out[k] = t + torch.normal(mean=0.0, std=sigma, size=t.shape)  # privacy.py:212,245
        else:
            out[k] = t + sigma * z[k]
    return out, sens, sigma


def apply_update_dp(w_local: Dict[str, torch.Tensor], w_global: Dict[str, torch.Tensor], epsilon: float,
                    delta: float, max_norm: float, z: Optional[Dict[str, torch.Tensor]] = None):
    """Client glue ``src/client/federated_trainer.py:428-469`` (not importable upstream: NameError at
    ``:262``): delta = local - global (``:437-443``) -> add_noise (``:447-451``) -> global + noisy (``:454-459``)."""
    d = {k: w_local[k] - w_global[k] for k in w_local}
    noisy, sens, sigma = add_noise(d, epsilon, delta, max_norm, z)
    return {k: w_global[k] + noisy[k] for k in w_global}, sens, sigma
