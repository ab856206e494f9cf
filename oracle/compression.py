"""Oracle (test infrastructure): update codecs restated in numpy.

Follows ``src/shared/compression.py`` (module not importable upstream without ``lz4``;
``oracle/make_golden.py`` stubs ``sys.modules['lz4']`` to run it):
  * ``QuantizationCompressor._quantize_tensor`` ``:203-228``: symmetric
    ``scale = 2*max|x| / (2^b - 1)``, ``zp = (2^b - 1) // 2``; asymmetric
    ``scale = (max - min)/(2^b - 1)``, ``zp = -round(min/scale)``;
    ``q = clamp(round_half_even(x/scale + zp), 0, 2^b - 1)`` stored uint8 (b<=8) / int16 / int32.
    x/scale and +zp happen in fp32 (tensor / python-float -> fp32 divide by fp32(scale)).
  * ``_dequantize_tensor`` ``:230-244``: ``(float(q) - zp) * scale`` in fp32.
  * ``TopKSparsificationCompressor._sparsify_tensor`` ``:327-344``: ``k = int(n*(1-sparsity))``,
    at least 1; indices of the k largest |x| (``torch.topk``), values gathered.
  * ``_desparsify_tensor`` ``:346-365``: scatter into zeros.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def quant_params(x: np.ndarray, bits: int = 8, symmetric: bool = True) -> Tuple[float, int]:
    levels = 2 ** max(1, min(32, bits))
    x = np.asarray(x, dtype=np.float32)
    if symmetric:
        max_val = float(np.abs(x).max())
        return (2 * max_val) / (levels - 1), (levels - 1) // 2
    mn, mx = float(x.min()), float(x.max())
    scale = (mx - mn) / (levels - 1)
    return scale, -round(mn / scale)


def quantize(x: np.ndarray, bits: int = 8, symmetric: bool = True):
    scale, zp = quant_params(x, bits, symmetric)
    levels = 2 ** max(1, min(32, bits))
    x = np.asarray(x, dtype=np.float32)
    v = x / np.float32(scale) + np.float32(zp)             # fp32 divide then fp32 add
    q = np.clip(np.rint(v), 0, levels - 1)                 # rint == round-half-even == torch.round
    dt = np.uint8 if bits <= 8 else (np.int16 if bits <= 16 else np.int32)
    return q.astype(dt), scale, zp


def dequantize(q: np.ndarray, scale: float, zp: int) -> np.ndarray:
    return (q.astype(np.float32) - np.float32(zp)) * np.float32(scale)


def topk_count(n: int, sparsity_ratio: float) -> int:
    s = max(0.0, min(1.0, sparsity_ratio))
    k = int(n * (1 - s))
    return k if k > 0 else 1


def sparsify(x: np.ndarray, sparsity_ratio: float = 0.9):
    """Returns (values[k], indices[k] int64) ordered by |x| descending (ties: lower index first,
    which is what a stable sort gives; torch.topk's tie order is unspecified so tests compare
    the reconstructed dense tensor and the index SET)."""
    flat = np.asarray(x, dtype=np.float32).reshape(-1)
    k = topk_count(flat.size, sparsity_ratio)
    order = np.argsort(-np.abs(flat), kind="stable")[:k]
    return flat[order], order.astype(np.int64)


def desparsify(values: np.ndarray, indices: np.ndarray, shape) -> np.ndarray:
    out = np.zeros(int(np.prod(shape)), dtype=np.float32)
    out[indices] = values
    return out.reshape(shape)
