"""Multi-GPU FedAvg over NVLink peer memory: host side of ``flb_fedavg_allreduce_p2p`` (include/flb.h, csrc/p2p_reduce.cu).

One process per GPU.  Every rank allocates one device region (flags, inbox, the global model row), exports it as a
cudaIpcMemHandle, the 64-byte handles travel through ``torch.distributed`` once at set-up, and every rank maps the peers'
regions.  After that a round's aggregation is ONE kernel per rank: the rank's weighted partial sum is pushed chunk by chunk
into the chunk owner's inbox while later chunks are still being summed, owners add the partial sums in rank order and store
the result into every rank's global row -- the reference's client -> coordinator upload plus the broadcast of the new
global model (src/aggregation/fedavg.py:56-124), with no NCCL call and no staging copy."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from .ops import _row_stride, as_weight_tensor


class _RawCuda:
    """CUDA array interface over a raw device pointer (zero-copy ``torch.as_tensor``)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class PeerFedAvg:
    """``PeerFedAvg(ld, device, rank, world, group)`` -- collective constructor (every rank of ``group`` must call it).

    ``global_row`` is a torch view of this rank's global model row inside the shared region; ``reduce(theta, w, P)`` leaves
    sum over ranks of sum_k w[k] * theta[k] there, bit-identical on every rank."""

    CHUNK = 2048            # floats per chunk at least: 8 KB messages, ~200 chunks for SimpleCNN, ~720 for CIFAR10CNN
    MAX_CHUNKS = 1024       # long rows get larger chunks instead of more of them: about one chunk per resident CTA and phase,
                            # so the system-scope fence + flag that closes a chunk is paid once per CTA, not tens of times

    def __init__(self, ld: int, device: torch.device, rank: int, world: int, group=None, chunk: Optional[int] = None):
        import torch.distributed as dist
        if world < 2 or world > 16:
            raise L.FlbError(f"PeerFedAvg: world size {world} not in 2..16")
        L.ensure_device(device)
        self.device, self.rank, self.world, self.group = device, rank, world, group
        self.layout = L.P2pLayout()
        if chunk is None:
            chunk = max(self.CHUNK, -(-int(ld) // self.MAX_CHUNKS))
            chunk = -(-chunk // 1024) * 1024
        nbytes = L.call_ll("flb_p2p_region_layout", int(ld), world, int(chunk), C.byref(self.layout))
        self._own = C.c_void_p()
        self._peers = [None] * world
        self.epoch = 0
        self._w_key, self._w_dev = None, None
        ok, err = 1, ""
        try:
            with torch.cuda.device(device):
                L.call("flb_p2p_alloc", C.byref(self._own), nbytes)
                handle = (C.c_ubyte * 64)()
                L.call("flb_p2p_export", self._own, handle)
        except L.FlbError as e:                      # every rank still has to take part in the exchange below
            ok, err, handle = 0, str(e), (C.c_ubyte * 64)()
        handles = [None] * world
        dist.all_gather_object(handles, (ok, bytes(handle)), group=group)
        if ok and all(h[0] for h in handles):
            try:
                with torch.cuda.device(device):
                    for r, (_, hb) in enumerate(handles):
                        if r == rank:
                            self._peers[r] = self._own.value
                            continue
                        p = C.c_void_p()
                        L.call("flb_p2p_open", (C.c_ubyte * 64).from_buffer_copy(hb), C.byref(p))
                        self._peers[r] = p.value
            except L.FlbError as e:
                ok, err = 0, str(e)
        else:
            ok = 0
        flags = [None] * world
        dist.all_gather_object(flags, ok, group=group)
        if not all(flags):
            self.close()
            raise L.FlbError(f"PeerFedAvg: peer memory unavailable on ranks {[r for r, f in enumerate(flags) if not f]} {err}")
        self._regions = (C.c_void_p * world)(*self._peers)
        self.global_row = torch.as_tensor(_RawCuda(self._own.value + self.layout.off_global, int(ld)), device=device)

    def reduce(self, theta: torch.Tensor, w, P: int) -> torch.Tensor:
        """All ranks call this once per round on the current stream; returns ``global_row[:P]``."""
        L.require_cuda_f32(theta, "theta")
        if theta.dim() == 1:
            theta = theta.unsqueeze(0)
        K, ld = theta.shape[0], _row_stride(theta)
        key = tuple(w) if not isinstance(w, torch.Tensor) else None     # the same client weights round after round: one upload
        if key is not None and key == self._w_key:
            wt = self._w_dev
        else:
            wt = as_weight_tensor(w, theta.device)
            self._w_key, self._w_dev = key, wt
        if wt.numel() != K:
            raise L.FlbError(f"PeerFedAvg.reduce: {wt.numel()} weights for {K} client rows")
        self.epoch += 1
        with torch.cuda.device(self.device):
            L.call("flb_fedavg_allreduce_p2p", L.ptr(theta), ld, L.ptr(wt), K, int(P), C.byref(self.layout), self._regions,
                   self.rank, self.epoch, L.stream_ptr(self.device))
        self._keep = (theta, wt)                     # alive until the next call: the kernel may still be queued
        return self.global_row[:P]

    def close(self) -> None:
        try:
            with torch.cuda.device(self.device):
                torch.cuda.synchronize(self.device)
                for r, p in enumerate(self._peers):
                    if p is not None and r != self.rank:
                        L.call("flb_p2p_close", C.c_void_p(p))
                if self._own.value:
                    L.call("flb_p2p_free", self._own)
        except Exception:
            pass
        self._peers = [None] * self.world
        self._own = C.c_void_p()

    def __del__(self):
        if getattr(self, "_own", None) is not None and self._own.value:
            self.close()
