"""Oracle (test infrastructure): one federated round on CPU, restating the client glue.

  * synthetic inputs: SURVEY.md section 8(d) / BASELINE.md section 3 (seeded, MNIST- / CIFAR-shaped)
  * client round:  ``src/client/federated_trainer.py:390-500`` (train -> delta -> add_noise ->
    global + noisy -> ModelUpdate(num_samples=samples_processed, training_loss=loss))
  * coordinator:   ``src/aggregation/fedavg.py:56-124``
Used by tests (as the checker) and by ``bench.py``'s cpu_baseline / ``--impl reference`` legs.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import fedavg as FA
from . import models as M
from . import privacy as PV
from . import training as TR

MNIST_SIZES = (480, 512, 544, 576)
CIFAR_SIZES = (416, 448, 480)


def client_num_samples(model: str, client_idx: int) -> int:
    sizes = MNIST_SIZES if model == "simple_cnn" else CIFAR_SIZES
    return sizes[client_idx % len(sizes)]


def synthetic_client_data(model: str, client_idx: int, n: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """x ~ N(0,1) fp32 [N_c, C, H, W], y ~ U{0..9} int64, generator seed 1000 + client_idx."""
    g = torch.Generator().manual_seed(1000 + client_idx)
    n = n or client_num_samples(model, client_idx)
    x = torch.randn((n,) + M.input_shape(model), generator=g, dtype=torch.float32)
    y = torch.randint(0, 10, (n,), generator=g, dtype=torch.int64)
    return x, y


def flatten(w: Dict[str, torch.Tensor], names: Sequence[str]) -> np.ndarray:
    return np.concatenate([w[n].detach().reshape(-1).numpy() for n in names]).astype(np.float32)


def unflatten(flat: np.ndarray, spec) -> Dict[str, torch.Tensor]:
    out, off = {}, 0
    for n, shp in spec.items():
        k = int(np.prod(shp))
        out[n] = torch.from_numpy(np.array(flat[off:off + k], dtype=np.float32)).reshape(shp)
        off += k
    return out


def client_round(model: str, w_global: Dict[str, torch.Tensor], x: torch.Tensor, y: torch.Tensor, *,
                 epochs: int = 1, lr: float = 1e-3, optimizer: str = "adam", batch_size: int = 32,
                 dp: bool = True, epsilon: float = 1.0, delta: float = 1e-5, max_norm: float = 1.0,
                 z: Optional[Dict[str, torch.Tensor]] = None, dropout_rate: float = 0.0,
                 bn_state: Optional[Dict[str, torch.Tensor]] = None):
    """federated_trainer.py:367-500 for one client.  Returns (weights to upload, loss, acc, samples)."""
    w = {k: v.clone() for k, v in w_global.items()}            # :378 set_model_weights(global)
    loss, acc, _, samples = TR.train_local_model(model, w, TR.make_batches(x, y, batch_size), epochs, lr,
                                                 optimizer, dropout_rate=dropout_rate, bn_state=bn_state)
    if dp:
        w, _, _ = PV.apply_update_dp(w, w_global, epsilon, delta, max_norm, z)
    return w, loss, acc, samples


def federated_round(model: str, w_global: Dict[str, torch.Tensor], num_clients: int, *, dp: bool = True,
                    zs: Optional[List[Dict[str, torch.Tensor]]] = None, data=None, **kw):
    """One full round: every client trains from ``w_global``; FedAvg by samples processed."""
    spec = M.model_spec(model)
    names = list(spec)
    thetas, ns, losses = [], [], []
    for c in range(num_clients):
        x, y = data[c] if data is not None else synthetic_client_data(model, c)
        bn = M.new_bn_state(model) if model == "cifar10_cnn" else None
        w, loss, _, n = client_round(model, w_global, x, y, dp=dp, z=None if zs is None else zs[c],
                                     bn_state=bn, **kw)
        thetas.append(flatten(w, names))
        ns.append(n)
        losses.append(loss)
    flat, avg_loss, idx, wts = FA.aggregate(np.stack(thetas), ns, losses, min_clients=min(2, num_clients))
    return unflatten(flat, spec), {"client_thetas": thetas, "num_samples": ns, "losses": losses,
                                   "avg_loss": avg_loss, "weights": wts}
