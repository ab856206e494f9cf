"""Tensor-level wrappers over the C ABI (flat ``[K, ld]`` client-major buffers in HBM).

Data layout: every client's model is one fp32 row of a ``[K, ld]`` matrix, ``ld`` = P rounded up to
32 floats so rows start 128 B aligned; the layer -> (offset, shape) table is ``layout.ParamLayout``.
All functions launch on the current torch stream and never synchronise."""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch

from . import _lib as L


def _row_stride(t: torch.Tensor) -> int:
    if t.dim() != 2 or t.stride(1) != 1:
        raise L.FlbError("expected a [K, ld] tensor with unit inner stride")
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def as_weight_tensor(w, device) -> torch.Tensor:
    """Host weights are Python float64 in the reference (fedavg.py:247-265) and become fp32 when they
    multiply an fp32 tensor; the same single rounding happens here."""
    if isinstance(w, torch.Tensor):
        return w.to(device=device, dtype=torch.float32).contiguous()
    return torch.tensor([float(x) for x in w], dtype=torch.float64).to(torch.float32).to(device)


def fedavg_weighted_sum(theta: torch.Tensor, w, P: Optional[int] = None, out: Optional[torch.Tensor] = None,
                        accumulate: bool = False) -> torch.Tensor:
    L.require_cuda_f32(theta, "theta")
    L.ensure_device(theta.device)
    K = theta.shape[0]
    ld = _row_stride(theta)
    P = theta.shape[1] if P is None else P
    wt = as_weight_tensor(w, theta.device)
    if wt.numel() != K:
        raise L.FlbError(f"fedavg_weighted_sum: {wt.numel()} weights for {K} client rows")
    if out is None:
        out = torch.empty(P, dtype=torch.float32, device=theta.device)
        accumulate = False
    if P == 0:
        return out
    with torch.cuda.device(theta.device):
        L.call("flb_fedavg_weighted_sum", L.ptr(theta), ld, L.ptr(wt), L.ptr(out), K, P, int(accumulate),
               L.stream_ptr(theta.device))
    return out


def fedavg_weighted_sum_ptrs(ptr_table: torch.Tensor, seg_off: torch.Tensor, w, K: int, Lyr: int, P: int,
                             device) -> torch.Tensor:
    L.ensure_device(device)
    wt = as_weight_tensor(w, device)
    out = torch.empty(P, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        L.call("flb_fedavg_weighted_sum_ptrs", L.ptr(ptr_table), L.ptr(seg_off), L.ptr(wt), L.ptr(out), K, Lyr, P,
               L.stream_ptr(device))
    return out


def fedavg_weighted_sum_q8(q: torch.Tensor, scale: torch.Tensor, zp: torch.Tensor, seg_off: torch.Tensor, w,
                           P: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    if not (q.is_cuda and q.dtype == torch.uint8):
        raise L.FlbError("fedavg_weighted_sum_q8: q must be a CUDA uint8 tensor")
    L.ensure_device(q.device)
    K, Lyr = q.shape[0], seg_off.numel() - 1
    wt = as_weight_tensor(w, q.device)
    if out is None:
        out = torch.empty(P, dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        L.call("flb_fedavg_weighted_sum_q8", L.ptr(q), _row_stride(q), L.ptr(scale), L.ptr(zp), L.ptr(seg_off),
               L.ptr(wt), L.ptr(out), K, Lyr, P, L.stream_ptr(q.device))
    return out


def gaussian_sigma_unit(epsilon: float, delta: float) -> float:
    """sigma / sensitivity of the reference's Gaussian mechanism (src/shared/privacy.py:202-209)."""
    if epsilon <= 0:
        raise ValueError("Epsilon must be positive")
    if delta <= 0 or delta >= 1:
        raise ValueError("Delta must be in (0, 1)")
    return math.sqrt(2 * math.log(1.25 / delta)) / epsilon


def dp_clip_noise(local: torch.Tensor, global_w: Optional[torch.Tensor], max_norm: float, sigma_unit: float,
                  seed: int = 0, stream_base: int = 0, z: Optional[torch.Tensor] = None, P: Optional[int] = None,
                  out: Optional[torch.Tensor] = None, stream_stride: int = 1, absmax_seg: Optional[torch.Tensor] = None):
    """Batched update-level DP over K client rows.  Returns (upload rows [K, ld], ||delta_k|| [K]); with ``absmax_seg`` (the
    layer offset table) also the per-(client, layer) max|upload| bit patterns [K, L] that ``q8_quantize(absmax=...)`` takes."""
    L.require_cuda_f32(local, "local")
    L.ensure_device(local.device)
    K, ld = local.shape[0], _row_stride(local)
    P = local.shape[1] if P is None else P
    dev = local.device
    if global_w is not None:
        L.require_cuda_f32(global_w, "global_w")
    if z is not None:
        L.require_cuda_f32(z, "z")
        if _row_stride(z) != ld:
            raise L.FlbError("dp_clip_noise: z must share the row pitch of local")
    if out is None:
        out = torch.empty_like(local)
    norm2 = torch.empty(K, dtype=torch.float64, device=dev)
    norms = torch.empty(K, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = L.stream_ptr(dev)
        L.call("flb_dp_sumsq", L.ptr(local), ld, L.ptr(global_w), L.ptr(norm2), K, P, st)
        if absmax_seg is not None:
            Lyr = absmax_seg.numel() - 1
            absmax = torch.empty((K, Lyr), dtype=torch.int32, device=dev)
            L.call("flb_dp_clip_noise_absmax", L.ptr(local), ld, L.ptr(global_w), L.ptr(z), L.ptr(norm2), L.ptr(out),
                   L.ptr(norms), float(max_norm), float(sigma_unit), int(seed) & (2**64 - 1),
                   int(stream_base) & (2**64 - 1), int(stream_stride), L.ptr(absmax_seg), Lyr, L.ptr(absmax), K, P, st)
            return out, norms, absmax
        L.call("flb_dp_clip_noise", L.ptr(local), ld, L.ptr(global_w), L.ptr(z), L.ptr(norm2), L.ptr(out),
               L.ptr(norms), float(max_norm), float(sigma_unit), int(seed) & (2**64 - 1),
               int(stream_base) & (2**64 - 1), int(stream_stride), K, P, st)
    return out, norms


def dp_add_noise(x: torch.Tensor, sigma: float, seed: int = 0, stream_base: int = 0,
                 z: Optional[torch.Tensor] = None, P: Optional[int] = None) -> torch.Tensor:
    L.require_cuda_f32(x, "x")
    L.ensure_device(x.device)
    K, ld = x.shape[0], _row_stride(x)
    P = x.shape[1] if P is None else P
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        L.call("flb_dp_add_noise", L.ptr(x), ld, L.ptr(z), L.ptr(out), float(sigma), int(seed) & (2**64 - 1),
               int(stream_base) & (2**64 - 1), K, P, L.stream_ptr(x.device))
    return out


def philox_normal(n: int, seed: int, stream: int, device) -> torch.Tensor:
    device = torch.device(device)
    L.ensure_device(device)
    out = torch.empty(n, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        L.call("flb_philox_normal", L.ptr(out), n, seed, stream, L.stream_ptr(device))
    return out


def philox_raw(nblocks: int, seed: int, stream: int, first_block: int, device) -> torch.Tensor:
    device = torch.device(device)
    L.ensure_device(device)
    out = torch.empty((nblocks, 4), dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        L.call("flb_philox_raw", L.ptr(out), nblocks, seed, stream, first_block, L.stream_ptr(device))
    return out


def q8_quantize(x: torch.Tensor, seg_off: torch.Tensor, P: Optional[int] = None, bits: int = 8,
                symmetric: bool = True, absmax: Optional[torch.Tensor] = None):
    """Per-(client, layer) affine quantisation of the K client rows.  Returns (q uint8 [K, ld], scale [K, L], zp [K, L]).
    ``absmax`` (from ``dp_clip_noise(absmax_seg=...)``, symmetric mode only) skips the max|x| reduction pass."""
    L.require_cuda_f32(x, "x")
    L.ensure_device(x.device)
    K, ld = x.shape[0], _row_stride(x)
    P = x.shape[1] if P is None else P
    Lyr = seg_off.numel() - 1
    dev = x.device
    q = torch.empty((K, ld), dtype=torch.uint8, device=dev)
    scale = torch.empty((K, Lyr), dtype=torch.float32, device=dev)
    zp = torch.empty((K, Lyr), dtype=torch.float32, device=dev)
    if absmax is not None:
        if not symmetric:
            raise L.FlbError("q8_quantize: a precomputed absmax only serves the symmetric mode")
        with torch.cuda.device(dev):
            L.call("flb_q8_quantize_absmax", L.ptr(x), ld, L.ptr(seg_off), L.ptr(absmax), L.ptr(q), ld, L.ptr(scale), L.ptr(zp),
                   K, Lyr, P, bits, L.stream_ptr(dev))
        return q, scale, zp
    scratch = torch.empty(2 * K * Lyr, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L.call("flb_q8_quantize", L.ptr(x), ld, L.ptr(seg_off), L.ptr(q), ld, L.ptr(scale), L.ptr(zp), L.ptr(scratch),
               K, Lyr, P, bits, int(symmetric), L.stream_ptr(dev))
    return q, scale, zp


def q8_dequantize(q: torch.Tensor, scale: torch.Tensor, zp: torch.Tensor, seg_off: torch.Tensor,
                  P: Optional[int] = None) -> torch.Tensor:
    L.ensure_device(q.device)
    K, ldq = q.shape[0], _row_stride(q)
    P = q.shape[1] if P is None else P
    Lyr = seg_off.numel() - 1
    out = torch.zeros((K, ldq), dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        L.call("flb_q8_dequantize", L.ptr(q), ldq, L.ptr(seg_off), L.ptr(scale), L.ptr(zp), L.ptr(out), ldq, K, Lyr, P,
               L.stream_ptr(q.device))
    return out


def topk_count(n: int, sparsity_ratio: float) -> int:
    """k = int(n * (1 - sparsity)), at least 1 (src/shared/compression.py:333-338)."""
    k = int(n * (1 - max(0.0, min(1.0, sparsity_ratio))))
    return k if k > 0 else 1


def topk_select(x: torch.Tensor, seg_off: torch.Tensor, kk: Sequence[int], P: Optional[int] = None):
    """Per-(client, layer) top-k by |x| over the K client rows.  Returns (idx int32 [K, ldk] layer-relative,
    val fp32 [K, ldk], out_off int64 [L+1] on the device, kk int32 [L] on the device)."""
    L.require_cuda_f32(x, "x")
    L.ensure_device(x.device)
    K, ld = x.shape[0], _row_stride(x)
    dev = x.device
    Lyr = seg_off.numel() - 1
    if len(kk) != Lyr:
        raise L.FlbError(f"topk_select: {len(kk)} counts for {Lyr} layers")
    offs = [0]
    for v in kk:
        offs.append(offs[-1] + int(v))
    ldk = max(offs[-1], 1)
    kk_t = torch.tensor([int(v) for v in kk], dtype=torch.int32, device=dev)
    off_t = torch.tensor(offs, dtype=torch.int64, device=dev)
    idx = torch.empty((K, ldk), dtype=torch.int32, device=dev)
    val = torch.empty((K, ldk), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.call("flb_topk_select", L.ptr(x), ld, L.ptr(seg_off), L.ptr(kk_t), L.ptr(off_t), L.ptr(idx), L.ptr(val), ldk, K, Lyr,
               L.stream_ptr(dev))
    return idx, val, off_t, kk_t


def topk_scatter(idx: torch.Tensor, val: torch.Tensor, seg_off: torch.Tensor, kk_t: torch.Tensor, off_t: torch.Tensor,
                 P: int, ld: Optional[int] = None) -> torch.Tensor:
    """Dense [K, ld] rows rebuilt from the (index, value) pairs (zeros elsewhere)."""
    dev = idx.device
    L.ensure_device(dev)
    K, ldk = idx.shape[0], _row_stride(idx)
    ld = ld or (P + 31) // 32 * 32
    out = torch.empty((K, ld), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.call("flb_topk_scatter", L.ptr(idx), L.ptr(val), ldk, L.ptr(seg_off), L.ptr(kk_t), L.ptr(off_t), L.ptr(out), ld, K,
               seg_off.numel() - 1, P, L.stream_ptr(dev))
    return out


def update_stats(ptr_table: torch.Tensor, seg_off: torch.Tensor, K: int, P: int, device):
    """One pass over K x L tensors addressed by a device pointer table: (max|x| fp32 [K, L], flags int32 [K, L] with
    bit 0 = NaN present, bit 1 = Inf present)."""
    L.ensure_device(device)
    Lyr = seg_off.numel() - 1
    mx = torch.empty((K, Lyr), dtype=torch.float32, device=device)
    fl = torch.empty((K, Lyr), dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        L.call("flb_update_stats", L.ptr(ptr_table), L.ptr(seg_off), L.ptr(mx), L.ptr(fl), K, Lyr, P, L.stream_ptr(device))
    return mx, fl


def rows_ptr_table(rows: torch.Tensor, offsets: Sequence[int]) -> torch.Tensor:
    """Pointer table of the layers of every row of a [K, ld] matrix (for update_stats / delta_norms)."""
    base, stride = rows.data_ptr(), _row_stride(rows) * 4
    return torch.tensor([base + k * stride + 4 * int(o) for k in range(rows.shape[0]) for o in offsets],
                        dtype=torch.int64).to(rows.device)


def delta_norms(new_ptrs: torch.Tensor, old_ptrs: torch.Tensor, seg_off: torch.Tensor, P: int, device) -> torch.Tensor:
    """Per layer (sum (new - old)^2, sum new^2) in double: [L, 2] on the device."""
    L.ensure_device(device)
    Lyr = seg_off.numel() - 1
    out = torch.empty((Lyr, 2), dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        L.call("flb_delta_norms", L.ptr(new_ptrs), L.ptr(old_ptrs), L.ptr(seg_off), L.ptr(out), Lyr, P, L.stream_ptr(device))
    return out
