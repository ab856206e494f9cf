"""Flat parameter layout: layer name -> (offset, shape) in reference ``named_parameters()`` order
(src/shared/models_pytorch.py:25-27), one fp32 row per client, row pitch padded to 32 floats."""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch


def _numel(shape) -> int:
    n = 1
    for s in shape:
        n *= int(s)
    return n


class ParamLayout:
    def __init__(self, spec: "OrderedDict[str, Tuple[int, ...]]"):
        self.names: List[str] = list(spec)
        self.shapes: Dict[str, Tuple[int, ...]] = {k: tuple(int(s) for s in v) for k, v in spec.items()}
        self.offsets: Dict[str, int] = {}
        off = 0
        for n in self.names:
            self.offsets[n] = off
            off += _numel(self.shapes[n])
        self.P = off
        self.ld = (off + 31) // 32 * 32
        self._seg_off_cache: Dict[str, torch.Tensor] = {}

    @classmethod
    def from_weights(cls, weights: Dict[str, torch.Tensor]) -> "ParamLayout":
        return cls(OrderedDict((k, tuple(v.shape)) for k, v in weights.items()))

    def signature(self):
        return tuple((n, self.shapes[n]) for n in self.names)

    def seg_off(self, device) -> torch.Tensor:
        key = str(device)
        if key not in self._seg_off_cache:
            offs = [self.offsets[n] for n in self.names] + [self.P]
            self._seg_off_cache[key] = torch.tensor(offs, dtype=torch.int64, device=device)
        return self._seg_off_cache[key]

    def new_rows(self, K: int, device, zero: bool = True) -> torch.Tensor:
        f = torch.zeros if zero else torch.empty
        return f((K, self.ld), dtype=torch.float32, device=device)

    def flatten_into(self, row: torch.Tensor, weights: Dict[str, torch.Tensor]) -> None:
        """Copy a dict of tensors (any device) into one [ld] row (H2D copies when the dict is on the host)."""
        for n in self.names:
            o = self.offsets[n]
            row[o:o + _numel(self.shapes[n])].copy_(weights[n].reshape(-1), non_blocking=True)

    def flatten(self, weights: Dict[str, torch.Tensor], device) -> torch.Tensor:
        row = torch.zeros(self.ld, dtype=torch.float32, device=device)
        self.flatten_into(row, weights)
        return row

    def views(self, row: torch.Tensor) -> "OrderedDict[str, torch.Tensor]":
        """Zero-copy dict of views into a flat row."""
        return OrderedDict((n, row[self.offsets[n]:self.offsets[n] + _numel(self.shapes[n])].view(self.shapes[n]))
                           for n in self.names)

    def unflatten(self, row: torch.Tensor, device=None) -> "OrderedDict[str, torch.Tensor]":
        """Fresh tensors (clones), optionally moved to ``device`` -- the ownership convention of
        ``get_model_weights`` (models_pytorch.py:25-27)."""
        out = OrderedDict()
        for n, v in self.views(row).items():
            out[n] = v.clone() if device is None else v.to(device, copy=True)
        return out
