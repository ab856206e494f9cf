"""Oracle (test infrastructure): restatement of the client hot loop.

Follows ``src/shared/training.py``:
  * ``LocalTrainer.train_local_model`` ``:60-171`` (fresh optimizer per call ``:89``,
    ``CrossEntropyLoss`` mean reduction ``:90``)
  * ``_train_epoch`` ``:173-212``  (zero_grad -> forward -> loss -> backward -> step;
    loss = mean over batches of the batch-mean loss ``:209``; accuracy = argmax matches / samples ``:210``)
  * ``_create_optimizer`` ``:244-255``: Adam(lr) | SGD(lr, momentum=0.9) | AdamW(lr)
The optimizer arithmetic is written out (torch.optim defaults: Adam betas (0.9, 0.999),
eps 1e-8, no weight decay; AdamW weight_decay 0.01 decoupled; SGD momentum 0.9, dampening 0,
first step buf = grad) so the CUDA optimizer kernel has a line-by-line counterpart.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import models as M


class OptState:
    def __init__(self, kind: str, lr: float, names: Sequence[str]):
        kind = kind.lower()
        if kind not in ("adam", "sgd", "adamw"):
            raise ValueError(f"Unknown optimizer type: {kind}")  # training.py:255
        self.kind, self.lr, self.t = kind, lr, 0
        self.m: Dict[str, torch.Tensor] = {}
        self.v: Dict[str, torch.Tensor] = {}

    def step(self, w: Dict[str, torch.Tensor], g: Dict[str, torch.Tensor]) -> None:
        self.t += 1
        lr = self.lr
        for n in w:
            p, grad = w[n], g[n]
            if self.kind == "sgd":
                # torch.optim.SGD(momentum=0.9): buf = grad (t=1) else 0.9*buf + grad; p -= lr*buf
                if n not in self.m:
                    self.m[n] = grad.clone()
                else:
                    self.m[n].mul_(0.9).add_(grad)
                p.add_(self.m[n], alpha=-lr)
                continue
            b1, b2, eps = 0.9, 0.999, 1e-8
            if self.kind == "adamw":
                p.mul_(1.0 - lr * 0.01)
            if n not in self.m:
                self.m[n] = torch.zeros_like(p)
                self.v[n] = torch.zeros_like(p)
            # torch.optim.Adam single-tensor path:
            #   m.lerp_(g, 1-b1); v = b2*v + (1-b2)*g*g
            #   denom = sqrt(v)/sqrt(1-b2^t) + eps ; p -= (lr/(1-b1^t)) * m/denom
            self.m[n].lerp_(grad, 1.0 - b1)
            self.v[n].mul_(b2).addcmul_(grad, grad, value=1.0 - b2)
            bc1 = 1.0 - b1 ** self.t
            bc2 = 1.0 - b2 ** self.t
            denom = (self.v[n].sqrt() / math.sqrt(bc2)).add_(eps)
            p.addcdiv_(self.m[n], denom, value=-(lr / bc1))


def loss_and_grads(model: str, w: Dict[str, torch.Tensor], x: torch.Tensor, y: torch.Tensor, **fw):
    """One forward/backward of the batch-mean cross-entropy (training.py:189-196)."""
    wl = {k: v.detach().clone().requires_grad_(True) for k, v in w.items()}
    logits = M.forward(model, wl, x, **fw)
    loss = F.cross_entropy(logits, y)
    grads = torch.autograd.grad(loss, list(wl.values()))
    return loss.detach(), logits.detach(), dict(zip(wl.keys(), grads))


def train_epoch(model: str, w: Dict[str, torch.Tensor], batches: Iterable[Tuple[torch.Tensor, torch.Tensor]],
                opt: OptState, dropout_rate: float = 0.0,
                masks_per_step: Optional[List[List[torch.Tensor]]] = None,
                bn_state: Optional[Dict[str, torch.Tensor]] = None):
    """training.py:173-212.  Mutates ``w`` (and ``bn_state``) in place."""
    running_loss, correct, total, nb = 0.0, 0, 0, 0
    for step, (x, y) in enumerate(batches):
        masks = masks_per_step[step] if masks_per_step is not None else None
        loss, logits, g = loss_and_grads(model, w, x, y, train=True, dropout_rate=dropout_rate,
                                         masks=masks, bn_state=bn_state)
        opt.step(w, g)
        running_loss += float(loss)
        correct += int((logits.argmax(1) == y).sum())
        total += int(y.numel())
        nb += 1
    return running_loss / max(nb, 1), correct / max(total, 1), total


def train_local_model(model: str, w: Dict[str, torch.Tensor], batches: Sequence, epochs: int,
                      learning_rate: float = 1e-3, optimizer_type: str = "adam", **kw):
    """training.py:60-171 without checkpoint / validation side paths.
    Returns (final_loss, final_accuracy, epochs_completed, samples_processed)."""
    opt = OptState(optimizer_type, learning_rate, list(w))
    loss = acc = 0.0
    total = 0
    done = 0
    for _ in range(epochs):
        loss, acc, n = train_epoch(model, w, batches, opt, **kw)
        total += n
        done += 1
    return loss, acc, done, total


def make_batches(x: torch.Tensor, y: torch.Tensor, batch_size: int = 32):
    """Pre-batched, unshuffled loader: any sized iterable of (data, targets)
    satisfies ``_train_epoch`` (training.py:184,209)."""
    return [(x[i:i + batch_size], y[i:i + batch_size]) for i in range(0, x.shape[0], batch_size)]
