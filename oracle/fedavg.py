"""Oracle (test infrastructure): FedAvg weighted aggregation, restated in numpy fp32.

Follows ``src/aggregation/fedavg.py``:
  * ``_calculate_sample_weights`` ``:247-256``  w_i = n_i / sum(n)   (Python float64)
  * ``_normalize_weights``        ``:258-265``
  * ``_weighted_average``         ``:267-289``  out = zeros_like(first); out += w_i * theta_i,
    sequentially in client order.  ``w_i * tensor`` with a Python float multiplies in the
    tensor dtype: the weight is rounded to fp32 first, the product is rounded to fp32, then
    the add is rounded to fp32 (no fused multiply-add).  The numpy statement below is therefore
    BIT-EXACT with the reference on CPU (checked in tests/test_oracle_golden.py).
  * ``aggregate_updates`` ``:56-124``: filtering (``:209-245``), min/max clients (``:78-86``),
    weighted loss (``:99-100``).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np


def sample_weights(num_samples: Sequence[int]) -> List[float]:
    total = sum(num_samples)
    if total == 0:
        return [1.0 / len(num_samples)] * len(num_samples)
    return [n / total for n in num_samples]


def normalize_weights(weights: Sequence[float]) -> List[float]:
    total = sum(weights)
    if total == 0:
        return [1.0 / len(weights)] * len(weights)
    return [w / total for w in weights]


def weighted_average_flat(theta: np.ndarray, weights: Sequence[float]) -> np.ndarray:
    """theta: [K, P] fp32 (client-major stack of flattened weights).  Sequential fp32 axpy."""
    theta = np.asarray(theta, dtype=np.float32)
    out = np.zeros(theta.shape[1], dtype=np.float32)
    for k, w in enumerate(weights):
        out = out + np.float32(w) * theta[k]          # fp32 mul, then fp32 add (two roundings)
    return out


def weighted_average(updates: List[Dict[str, np.ndarray]], weights: Sequence[float]) -> Dict[str, np.ndarray]:
    out = {name: np.zeros_like(np.asarray(t, dtype=np.float32)) for name, t in updates[0].items()}
    for upd, w in zip(updates, weights):
        for name, t in upd.items():
            if name in out:
                out[name] = out[name] + np.float32(w) * np.asarray(t, dtype=np.float32)
    return out


def select_clients(num_samples: Sequence[int], losses: Sequence[float], min_clients: int = 2,
                   max_clients: Optional[int] = None) -> List[int]:
    """Indices that survive ``_filter_and_validate_updates`` basic checks (``:216-222``) and the
    max_clients truncation (stable sort by num_samples descending, ``:82-85``)."""
    idx = [i for i, (n, l) in enumerate(zip(num_samples, losses)) if n > 0 and l >= 0]
    if len(idx) < min_clients:
        raise ValueError(f"Insufficient valid updates: {len(idx)} < {min_clients}")
    if max_clients and len(idx) > max_clients:
        idx = sorted(idx, key=lambda i: num_samples[i], reverse=True)[:max_clients]
    return idx


def aggregate(theta: np.ndarray, num_samples: Sequence[int], losses: Sequence[float],
              weights: Optional[Sequence[float]] = None, min_clients: int = 2,
              max_clients: Optional[int] = None):
    """``aggregate_updates`` on a flat [K, P] stack.  Returns (theta_global[P], avg_loss, kept indices, weights)."""
    idx = select_clients(num_samples, losses, min_clients, max_clients)
    if weights is None:
        w = sample_weights([num_samples[i] for i in idx])
    else:
        w = normalize_weights(list(weights)[:len(idx)])     # fedavg.py:92 (positional slice, as upstream)
    out = weighted_average_flat(np.asarray(theta)[idx], w)
    avg_loss = sum(losses[i] * wi for i, wi in zip(idx, w))
    return out, avg_loss, idx, w
