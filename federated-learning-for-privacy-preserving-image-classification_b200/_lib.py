"""ctypes binding of libflb.so (the C ABI declared in include/flb.h).

There is deliberately no fallback: if the shared library is missing, or no sm_100 device is
visible, every compute entry raises ``FlbError`` instead of silently running PyTorch ops."""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

import torch

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libflb.so")


def fresh_seed() -> int:
    """64 bits of OS entropy: the default Philox seed of every object that draws DP noise or dropout masks (the reference
    draws from torch's nondeterministically seeded global RNG; a fixed default would make the noise reproducible by anyone)."""
    return int.from_bytes(os.urandom(8), "little")


class FlbError(RuntimeError):
    """A libflb.so call failed (or the library / device is unavailable)."""


_vp, _i, _ll, _ull, _d = C.c_void_p, C.c_int, C.c_longlong, C.c_ulonglong, C.c_double

# name -> argtypes; mirrors include/flb.h one-to-one (tests/test_cabi.py checks both directions)
SIGNATURES = {
    "flb_version": [],
    "flb_init": [_i],
    "flb_l2_persist_window": [_vp, _ll, _vp],
    "flb_fedavg_weighted_sum": [_vp, _ll, _vp, _vp, _i, _ll, _i, _vp],
    "flb_fedavg_weighted_sum_ptrs": [_vp, _vp, _vp, _vp, _i, _i, _ll, _vp],
    "flb_fedavg_weighted_sum_q8": [_vp, _ll, _vp, _vp, _vp, _vp, _vp, _i, _i, _ll, _vp],
    "flb_p2p_region_layout": [_ll, _i, _i, _vp],
    "flb_p2p_alloc": [_vp, _ll],
    "flb_p2p_free": [_vp],
    "flb_p2p_export": [_vp, _vp],
    "flb_p2p_open": [_vp, _vp],
    "flb_p2p_close": [_vp],
    "flb_fedavg_allreduce_p2p": [_vp, _ll, _vp, _i, _ll, _vp, _vp, _i, C.c_uint, _vp],
    "flb_dp_sumsq": [_vp, _ll, _vp, _vp, _i, _ll, _vp],
    "flb_dp_clip_noise": [_vp, _ll, _vp, _vp, _vp, _vp, _vp, _d, _d, _ull, _ull, _ull, _i, _ll, _vp],
    "flb_dp_clip_noise_absmax": [_vp, _ll, _vp, _vp, _vp, _vp, _vp, _d, _d, _ull, _ull, _ull, _vp, _i, _vp, _i, _ll, _vp],
    "flb_dp_add_noise": [_vp, _ll, _vp, _vp, _d, _ull, _ull, _i, _ll, _vp],
    "flb_philox_normal": [_vp, _ll, _ull, _ull, _vp],
    "flb_philox_raw": [_vp, _ll, _ull, _ull, _ull, _vp],
    "flb_q8_quantize": [_vp, _ll, _vp, _vp, _ll, _vp, _vp, _vp, _i, _i, _ll, _i, _i, _vp],
    "flb_q8_quantize_absmax": [_vp, _ll, _vp, _vp, _vp, _ll, _vp, _vp, _i, _i, _ll, _i, _vp],
    "flb_q8_dequantize": [_vp, _ll, _vp, _vp, _vp, _vp, _ll, _i, _i, _ll, _vp],
    "flb_update_stats": [_vp, _vp, _vp, _vp, _i, _i, _ll, _vp],
    "flb_delta_norms": [_vp, _vp, _vp, _vp, _i, _ll, _vp],
    "flb_gather_normalize_u8": [_vp, _ll, _i, _i, _i, _i, _vp, _ll, _vp, _vp, _vp, _vp],
    "flb_gather_labels": [_vp, _vp, _vp, _ll, _vp],
    "flb_topk_select": [_vp, _ll, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _vp],
    "flb_topk_scatter": [_vp, _vp, _ll, _vp, _vp, _vp, _vp, _ll, _i, _i, _ll, _vp],
    "flb_train_ws_bytes": [_i, _i, _i],
    "flb_train_ws_offset": [_i, _i, _i, C.c_char_p],
    "flb_train_bn_floats": [_i],
    "flb_train_begin_epoch": [_vp, _vp],
    "flb_train_begin_round": [_vp, _vp, _vp],
    "flb_train_step": [_vp, _vp],
    "flb_train_step_grads": [_vp, _vp, _ll, _vp],
    "flb_train_forward_backward": [_vp, _vp],
    "flb_train_forward": [_vp, _vp],
    "flb_train_advance": [_vp, _vp],
    "flb_train_step_launches": [_vp],
    "flb_debug_trace_set": [_vp, _i],
    "flb_mma_microbench": [_i, _i, _i, _i, _i, _i, _vp, _vp],
    "flb_train_step_profiled": [_vp, _vp, C.c_char_p, _i, _vp, _i],
}
_RESTYPE_LL = {"flb_train_ws_bytes", "flb_train_ws_offset", "flb_train_bn_floats", "flb_p2p_region_layout", "flb_l2_persist_window"}


class P2pLayout(C.Structure):
    """``flb_p2p_layout`` of include/flb.h, field for field."""
    _fields_ = [("bytes", _ll), ("off_flags_a", _ll), ("off_flags_b", _ll), ("off_inbox", _ll), ("off_global", _ll), ("ld", _ll),
                ("world", _i), ("chunk", _i)]



class TrainArgs(C.Structure):
    """``flb_train_args`` of include/flb.h, field for field."""
    _fields_ = [
        ("x", _vp), ("y", _vp), ("sample_off", _vp), ("nsamples", _vp), ("step_ctr", _vp),
        ("W", _vp), ("G", _vp), ("M", _vp), ("V", _vp), ("tcount", _vp), ("ws", _vp),
        ("loss_sum", _vp), ("correct", _vp), ("nbatch", _vp), ("nseen", _vp),
        ("drop_keep", _vp), ("dp_z", _vp), ("bn_running", _vp), ("epoch_nonce", _vp),
        ("ld", _ll), ("seed", _ull), ("client_base", _ull), ("client_stride", _ull),
        ("lr", _d), ("beta1", _d), ("beta2", _d), ("eps", _d), ("weight_decay", _d), ("momentum", _d),
        ("model", _i), ("K", _i), ("B", _i), ("precision", _i), ("opt", _i), ("dp_mode", _i), ("eval_mode", _i), ("tc_mask", _i),
        ("drop_p", C.c_float), ("dp_clip", C.c_float), ("dp_sigma", C.c_float),
    ]


_lock = threading.Lock()
_lib: Optional[C.CDLL] = None
_inited = set()


def load() -> C.CDLL:
    """dlopen libflb.so and type its entry points.  Needs no GPU (symbol check only)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise FlbError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        lib.flb_last_error.restype = C.c_char_p
        lib.flb_last_error.argtypes = []
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here == header/library drift
            fn.restype = C.c_longlong if name in _RESTYPE_LL else C.c_int
            fn.argtypes = argtypes
        _lib = lib
        return lib


def ensure_device(device: torch.device) -> None:
    if device.type != "cuda":
        raise FlbError(f"flb200 kernels need a CUDA (sm_100a) device, got '{device}'; there is no CPU fallback")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx in _inited:
        return
    lib = load()
    rc = lib.flb_init(idx)
    if rc != 0:
        raise FlbError(lib.flb_last_error().decode())
    _inited.add(idx)


def call(name: str, *args) -> None:
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise FlbError(f"{name}: {lib.flb_last_error().decode()} (code {rc})")


def call_ll(name: str, *args) -> int:
    v = getattr(load(), name)(*args)
    if v < 0:
        raise FlbError(f"{name}{args}: unsupported configuration")
    return v


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device: Optional[torch.device] = None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda_f32(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise FlbError(f"{what}: expected a CUDA tensor (no CPU fallback)")
    if t.dtype != torch.float32:
        raise FlbError(f"{what}: expected float32, got {t.dtype}")
    return t
