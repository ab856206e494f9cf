"""Local client training behind the reference's ``LocalTrainer`` surface, plus the batched engine under it.

``BatchedClientTrainer`` keeps K clients resident on one GPU as rows of ``[K, ld]`` parameter / gradient /
optimizer-moment matrices and advances all of them one minibatch per launch sequence (``flb_train_step``);
a whole epoch is captured once in a CUDA graph and replayed, and the loss / accuracy accumulators
(``running_loss``, ``correct_predictions``, src/shared/training.py:200-203) stay on the device until the epoch ends --
one host read per epoch instead of two syncs per step.

``LocalTrainer`` mirrors src/shared/training.py:28-403 (constructor, ``train_local_model``, ``evaluate_model``,
gradient hooks, checkpoints, ``TrainingError``) and is a K = 1 view of the same engine."""
from __future__ import annotations

import ctypes as C
import json
import logging
import math
import os
import time
from collections import OrderedDict
from dataclasses import dataclass, field
from datetime import datetime
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib as L
from .layout import ParamLayout
from .models import TrainingMetrics
from .models_pytorch import INPUT_SHAPES, MODEL_IDS, FederatedCNNBase, ModelFactory

logger = logging.getLogger(__name__)

OPTIMIZERS = {"adam": 0, "sgd": 1, "adamw": 2}
PRECISIONS = {"fp32": 0, "tf32": 1}
MAX_BATCH = 32


class TrainingError(Exception):
    pass


def model_layout(model_name: str, num_classes: int = 10) -> ParamLayout:
    m = ModelFactory.create_model(model_name, num_classes=num_classes)
    return ParamLayout(m.param_spec())


def _cuda_device(device) -> torch.device:
    if device is None:
        return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cuda")
    device = torch.device(device)
    if device.type == "cuda" and device.index is None and torch.cuda.is_available():
        device = torch.device("cuda", torch.cuda.current_device())
    return device


class BatchedClientTrainer:
    """K simulated clients training side by side on one GPU (the "batched across clients" engine)."""

    def __init__(self, model_name: str, num_clients: int, device=None, batch_size: int = 32,
                 dropout_rate: Optional[float] = None, precision: str = "fp32", seed: Optional[int] = None,
                 client_base: int = 0, client_stride: int = 1, use_graph: bool = True):
        if model_name not in MODEL_IDS:
            raise ValueError(f"Unknown model: {model_name}. Available: {list(MODEL_IDS)}")
        if not (1 <= batch_size <= MAX_BATCH):
            raise L.FlbError(f"batch_size must be in 1..{MAX_BATCH}, got {batch_size}")
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {list(PRECISIONS)}")
        self.model_name = model_name
        self.model_id = MODEL_IDS[model_name]
        self.K = int(num_clients)
        self.B = int(batch_size)
        self.device = _cuda_device(device)
        L.ensure_device(self.device)
        self.precision = precision
        # Philox seed of dropout masks and per-sample-DP noise.  Default: fresh entropy per engine (torch's global RNG is
        # nondeterministically seeded upstream too); pass ``seed=`` only for reproducible tests.
        self.seed = L.fresh_seed() if seed is None else int(seed) & (2**64 - 1)
        self.client_base = client_base
        self.client_stride = client_stride
        self.use_graph = use_graph
        self.dropout_rate = (0.25 if model_name == "simple_cnn" else 0.3) if dropout_rate is None else float(dropout_rate)
        self.layout = model_layout(model_name)
        self.sample_numel = math.prod(INPUT_SHAPES[model_name])
        dev, K = self.device, self.K
        lay = self.layout
        # gradients | parameters | Adam moments back to back in one allocation: ONE L2 access-policy window can cover a suffix
        # of it (see _launch_epoch).  FLB_L2_WINDOW = mv (default) | wmv | gwmv selects the suffix, FLB_NO_L2_PERSIST=1 none.
        self._state = torch.zeros((4, K, lay.ld), dtype=torch.float32, device=dev)
        self.G, self.W, self.M, self.V = self._state[0], self._state[1], self._state[2], self._state[3]
        first = {"mv": 2, "wmv": 1, "gwmv": 0}[os.environ.get("FLB_L2_WINDOW", "mv")]
        self._MV = self._state[first:]
        self.l2_persist = os.environ.get("FLB_NO_L2_PERSIST") is None
        self.tcount = torch.zeros(K, dtype=torch.int32, device=dev)
        self.step_ctr = torch.zeros(2, dtype=torch.int32, device=dev)
        self.epoch_nonce = torch.zeros(1, dtype=torch.int64, device=dev)   # +1 per epoch on the device, never reset (flb.h)
        # epoch accumulators in ONE buffer ([4, K] 32-bit words) so that they reach the host in one copy
        self._metrics = torch.zeros((4, K), dtype=torch.int32, device=dev)
        self.loss_sum = self._metrics[0].view(torch.float32)
        self.correct, self.nbatch, self.nseen = self._metrics[1], self._metrics[2], self._metrics[3]
        self._metrics_host = torch.zeros((4, K), dtype=torch.int32).pin_memory()
        self.ws_bytes = L.call_ll("flb_train_ws_bytes", self.model_id, K, self.B)
        self.ws = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=dev)      # pads of the NHWC grids stay zero forever
        nbn = L.call_ll("flb_train_bn_floats", self.model_id)
        self.bn_running = None                # cifar10_cnn: client-local running_mean | running_var, [K, 2 * 448]
        if nbn:
            self.bn_running = torch.zeros((K, nbn), dtype=torch.float32, device=dev)
            self.bn_running[:, nbn // 2:] = 1.0
        self.x = self.y = self.sample_off = self.nsamples = None
        self.n_host: List[int] = []
        self.drop_keep: Optional[torch.Tensor] = None
        self.dp_z: Optional[torch.Tensor] = None
        self.dp_mode, self.dp_clip, self.dp_sigma = 0, 1.0, 0.0
        self.tc_mask = 0                      # 0 = every GEMM on tensor cores when precision == 'tf32' (see flb.h)
        self._graphs: Dict[Any, Any] = {}        # captured epochs, keyed by the argument block (a few: one per sample buffer)
        self._seen_keys = set()
        self._staged = None                      # (x, y, event) uploaded ahead of time by prefetch_packed
        self._spare = None                       # the sample buffers not in use (double buffering)
        self._copy_stream = None
        self.keep_last_grads = False              # LocalTrainer: the epoch's last step also stores its gradient rows
        self.last_grads: Optional[torch.Tensor] = None
        self.args = L.TrainArgs()

    # ---- state ---------------------------------------------------------------------------------------
    def set_global_row(self, row: torch.Tensor) -> None:
        """Every client starts the round from the same global model (src/client/federated_trainer.py:378)."""
        self.W.copy_(row.to(self.device).view(1, -1).expand(self.K, -1))

    def begin_round(self, row: torch.Tensor) -> None:
        """Global model into every client row + fresh optimizer state (moments and step counts zero), one launch."""
        if self.x is None or row.device != self.W.device or row.numel() < self.layout.ld or row.data_ptr() % 16:
            self.set_global_row(row)
            self.M.zero_(); self.V.zero_(); self.tcount.zero_()
            return
        self._fill_args(0.0, "adam", train=True)
        with torch.cuda.device(self.device):
            L.call("flb_train_begin_round", C.byref(self.args), L.ptr(row), L.stream_ptr(self.device))

    def set_client_weights(self, k: int, weights: Dict[str, torch.Tensor]) -> None:
        self.layout.flatten_into(self.W[k], weights)

    def client_weights(self, k: int, device=None) -> Dict[str, torch.Tensor]:
        return self.layout.unflatten(self.W[k], device)

    def load_data(self, xs: Sequence[torch.Tensor], ys: Sequence[torch.Tensor]) -> None:
        """xs[k]: [N_k, C, H, W] fp32, ys[k]: [N_k] integer labels (host or device); batches are consecutive slices
        of ``batch_size`` samples, the last one ragged -- exactly what an unshuffled DataLoader yields."""
        if len(xs) != self.K or len(ys) != self.K:
            raise L.FlbError(f"load_data: expected {self.K} clients, got {len(xs)}")
        ns = [int(x.shape[0]) for x in xs]
        for x, y in zip(xs, ys):
            if x[0].numel() != self.sample_numel if x.shape[0] else False:
                raise L.FlbError(f"load_data: sample shape {tuple(x.shape[1:])} does not match {self.model_name}")
            if y.shape[0] != x.shape[0]:
                raise L.FlbError("load_data: data / target length mismatch")
        self._reserve(ns)
        off = 0
        for x, y, n in zip(xs, ys, ns):
            if n:
                self.x[off:off + n].copy_(x.reshape(n, -1), non_blocking=True)
                self.y[off:off + n].copy_(y, non_blocking=True)          # dtype conversion (int64 -> int32) on the device
            off += n

    def _reserve(self, ns: Sequence[int]) -> None:
        """Device sample store for these per-client counts; buffers (and therefore the captured graph) are reused
        while the counts stay the same."""
        ns = [int(n) for n in ns]
        dev, total = self.device, sum(ns)
        if self.x is None or self.x.shape[0] != max(total, 1):
            self.x = torch.empty((max(total, 1), self.sample_numel), dtype=torch.float32, device=dev)
            self.y = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        self._reserve_meta(ns)

    def _reserve_meta(self, ns: Sequence[int]) -> None:
        ns = [int(n) for n in ns]
        if ns != self.n_host:
            offs = [0]
            for n in ns[:-1]:
                offs.append(offs[-1] + n)
            self.sample_off = torch.tensor(offs, dtype=torch.int64).to(self.device)
            self.nsamples = torch.tensor(ns, dtype=torch.int32).to(self.device)
            self.n_host = ns
        self.h2d_bytes = sum(ns) * (self.sample_numel * 4 + 4)

    def attach(self, x_dev: torch.Tensor, y_dev: torch.Tensor, ns: Sequence[int]) -> None:
        """Use an already device-resident packed sample store (``DeviceShardBuilder.build``) in place: no copy."""
        if len(ns) != self.K:
            raise L.FlbError(f"attach: expected {self.K} clients, got {len(ns)}")
        total = sum(int(n) for n in ns)
        if not (x_dev.is_cuda and y_dev.is_cuda and x_dev.dtype == torch.float32 and y_dev.dtype == torch.int32
                and x_dev.is_contiguous() and y_dev.is_contiguous()):
            raise L.FlbError("attach: need contiguous CUDA tensors, x fp32 [sum N, C*H*W] and y int32 [sum N]")
        if x_dev.shape[0] < max(total, 1) or x_dev[0].numel() != self.sample_numel or y_dev.shape[0] < max(total, 1):
            raise L.FlbError("attach: x / y do not match the per-client counts or the model's sample shape")
        self.x = self.y = None
        self.n_host = []
        self._spare = self._staged = None
        self._reserve_meta(ns)
        self.x, self.y = x_dev, y_dev

    def prefetch_packed(self, x_all: torch.Tensor, y_all: torch.Tensor, ns: Sequence[int]) -> None:
        """Upload the NEXT round's samples (same per-client counts as the current ones) into the spare device buffers on
        a copy stream, so the host -> device transfer overlaps the round that is running; ``use_prefetched()`` swaps
        them in.  The copy waits only for the spare buffers' last reader (the work queued before the swap that freed them),
        not for a round enqueued since -- so it may be issued right after ``start_round`` and still overlap that round."""
        ns = [int(n) for n in ns]
        if self.x is None or ns != self.n_host:
            raise L.FlbError("prefetch_packed: call load_packed once with these per-client counts first")
        total = sum(ns)
        if x_all.shape[0] != total or y_all.shape[0] != total or x_all.numel() != total * self.sample_numel:
            raise L.FlbError("prefetch_packed: x_all / y_all do not match the per-client counts")
        with torch.cuda.device(self.device):
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(self.device)
            free_ev = getattr(self, "_spare_free_ev", None)
            if self._spare is None or self._spare[0].shape != self.x.shape:
                self._spare = (torch.empty_like(self.x), torch.empty_like(self.y))
                free_ev = None
            xs, ys = self._spare
            if free_ev is not None:
                self._copy_stream.wait_event(free_ev)
            else:
                self._copy_stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self._copy_stream):
                xs[:total].copy_(x_all.reshape(total, -1), non_blocking=True)
                ys[:total].copy_(y_all, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            self._staged = (xs, ys, ev)
            self._spare = None

    def use_prefetched(self) -> None:
        """Make the prefetched samples current (the compute stream waits for their copy); the old buffers become spare."""
        if self._staged is None:
            raise L.FlbError("use_prefetched: nothing was prefetched")
        xs, ys, ev = self._staged
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        self._spare_free_ev = torch.cuda.Event()          # everything that read the old buffers is queued before this point
        self._spare_free_ev.record(cur)
        self._spare, self._staged = (self.x, self.y), None
        self.x, self.y = xs, ys

    def load_packed(self, x_all: torch.Tensor, y_all: torch.Tensor, ns: Sequence[int]) -> None:
        """Same as ``load_data`` for inputs that are already concatenated client after client: ``x_all`` [sum N, C*H*W]
        (or [sum N, C, H, W]) fp32 and ``y_all`` [sum N] int32, ideally in pinned host memory -- two asynchronous
        host -> device copies for the whole round instead of two per client."""
        if len(ns) != self.K:
            raise L.FlbError(f"load_packed: expected {self.K} clients, got {len(ns)}")
        total = sum(int(n) for n in ns)
        if x_all.shape[0] != total or y_all.shape[0] != total or x_all.numel() != total * self.sample_numel:
            raise L.FlbError("load_packed: x_all / y_all do not match the per-client counts")
        self._reserve(ns)
        if total:
            self.x[:total].copy_(x_all.reshape(total, -1), non_blocking=True)
            self.y[:total].copy_(y_all, non_blocking=True)

    def configure_dp(self, mode: str = "none", max_grad_norm: float = 1.0, sigma: float = 0.0,
                     z: Optional[torch.Tensor] = None) -> None:
        """mode 'none' (reference behaviour: plain minibatch descent) or 'per_sample' (north-star kernel 2):
        g = (sum_i clip_C(g_i) + N(0, sigma^2)) / B inside every step.  ``z`` [K, ld] injects the standard normals."""
        if mode not in ("none", "per_sample"):
            raise ValueError("dp mode must be 'none' or 'per_sample'")
        self.dp_mode = 0 if mode == "none" else 1
        self.dp_clip, self.dp_sigma, self.dp_z = float(max_grad_norm), float(sigma), z
        self._graphs.clear()

    # ---- launches --------------------------------------------------------------------------------------
    def _fill_args(self, lr: float, optimizer_type: str, train: bool = True) -> None:
        opt = optimizer_type.lower()
        if opt not in OPTIMIZERS:
            raise ValueError(f"Unknown optimizer type: {optimizer_type}")          # training.py:255
        if self.x is None:
            raise L.FlbError("no client data loaded")
        a = self.args
        p = lambda t: None if t is None else t.data_ptr()      # noqa: E731
        a.x, a.y, a.sample_off, a.nsamples, a.step_ctr = p(self.x), p(self.y), p(self.sample_off), p(self.nsamples), p(self.step_ctr)
        a.W, a.G, a.M, a.V, a.tcount, a.ws = p(self.W), p(self.G), p(self.M), p(self.V), p(self.tcount), p(self.ws)
        a.loss_sum, a.correct, a.nbatch, a.nseen = p(self.loss_sum), p(self.correct), p(self.nbatch), p(self.nseen)
        a.drop_keep, a.dp_z, a.bn_running = p(self.drop_keep), p(self.dp_z), p(self.bn_running)
        a.epoch_nonce = p(self.epoch_nonce)
        a.eval_mode = 0 if train else 1
        a.ld, a.seed, a.client_base, a.client_stride = self.layout.ld, self.seed, self.client_base, self.client_stride
        a.lr, a.beta1, a.beta2, a.eps = float(lr), 0.9, 0.999, 1e-8            # torch.optim.Adam / AdamW defaults
        a.weight_decay = 0.01 if opt == "adamw" else 0.0
        a.momentum = 0.9                                                          # training.py:251
        a.model, a.K, a.B = self.model_id, self.K, self.B
        a.precision, a.opt, a.dp_mode = PRECISIONS[self.precision], OPTIMIZERS[opt], self.dp_mode if train else 0
        a.tc_mask = self.tc_mask
        a.drop_p = self.dropout_rate if train else 0.0
        a.dp_clip, a.dp_sigma = self.dp_clip, self.dp_sigma

    def max_steps(self) -> int:
        return (max(self.n_host) + self.B - 1) // self.B if self.n_host else 0

    def _launch_epoch(self) -> None:
        st = L.stream_ptr(self.device)
        ap = C.byref(self.args)
        # The Adam moments are touched once per step by the optimizer kernel; when they fit the L2 set-aside (about 10 SimpleCNN
        # clients) they are pinned there for the epoch instead of being evicted by the activations in between.  The window is a
        # stream attribute (recorded into the kernel nodes when this runs under capture); larger states get no window at all.
        if self.l2_persist:
            L.load().flb_l2_persist_window(L.ptr(self._MV), self._MV.numel() * 4, st)
        L.call("flb_train_begin_epoch", ap, st)
        n = self.max_steps()
        for s in range(n):
            if self.keep_last_grads and s == n - 1:
                L.call("flb_train_step_grads", ap, L.ptr(self.last_grads), self.layout.ld, st)
            else:
                L.call("flb_train_step", ap, st)

    def _run_epoch(self) -> None:
        """Eager the first time a configuration is seen, captured into one CUDA graph (begin_epoch + every step of the
        epoch) the second time, replayed afterwards.  All per-step variation is device-side, so the graph is static."""
        with torch.cuda.device(self.device):
            if self.keep_last_grads and self.last_grads is None:
                self.last_grads = self.layout.new_rows(self.K, self.device)      # before any capture: no allocation inside
            if self.l2_persist:               # sizes (or gives back) the device's L2 set-aside for THIS engine, replays included
                L.load().flb_l2_persist_window(L.ptr(self._MV), self._MV.numel() * 4, L.stream_ptr(self.device))
            key = (bytes(self.args), self.max_steps(), self.keep_last_grads)
            if self.use_graph and key in self._graphs:
                self._graphs[key].replay()
                return
            if self.use_graph and key in self._seen_keys:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):            # capture only records; nothing executes here
                    self._launch_epoch()
                if len(self._graphs) >= 4:
                    self._graphs.clear()
                self._graphs[key] = g
                g.replay()
                return
            self._seen_keys.add(key)
            self._launch_epoch()

    @property
    def _graph(self):
        """Any captured epoch graph (tests check that replay happened)."""
        return next(iter(self._graphs.values()), None)

    def profile_step(self) -> "OrderedDict[str, float]":
        """Per-kernel device milliseconds of ONE step (CUDA events after every kernel; synchronises).  The step is a
        real one: it advances the epoch's minibatch counter."""
        buf = C.create_string_buffer(4096)
        ms = (C.c_float * 48)()
        with torch.cuda.device(self.device):
            n = L.load().flb_train_step_profiled(C.byref(self.args), L.stream_ptr(self.device), buf, 4096, ms, 48)
        if n < 0:
            raise L.FlbError(L.load().flb_last_error().decode())
        names = buf.value.decode().split("\n")[:n]
        out: "OrderedDict[str, float]" = OrderedDict()
        for nm, v in zip(names, ms):
            out[nm] = out.get(nm, 0.0) + float(v)
        return out

    def launches_per_epoch(self) -> int:
        prologue = 2 if self.precision == "tf32" else 1           # begin_epoch (+ the tap-major weight repack)
        return prologue + self.max_steps() * int(L.load().flb_train_step_launches(C.byref(self.args)))

    def train(self, epochs: int, learning_rate: float = 0.001, optimizer_type: str = "adam"):
        """One call = one ``train_local_model`` for every client: fresh optimizer state (training.py:89), ``epochs``
        passes over each client's samples.  Returns per-client lists (loss, accuracy, samples_processed) where loss /
        accuracy are those of the last epoch (training.py:143-144) and samples are summed over epochs (:105)."""
        self._fill_args(learning_rate, optimizer_type, train=True)
        self.M.zero_()
        self.V.zero_()
        self.tcount.zero_()
        loss = acc = None
        total = torch.zeros(self.K, dtype=torch.int64)
        for _ in range(epochs):
            self._run_epoch()
            loss, acc, seen = self.epoch_metrics()
            total += seen
        if loss is None:
            return [0.0] * self.K, [0.0] * self.K, [0] * self.K
        return loss.tolist(), acc.tolist(), total.tolist()

    def epoch_metrics(self):
        """The one device->host read of an epoch: mean of batch-mean losses, accuracy, samples (training.py:209-212)."""
        self.enqueue_metrics_read()
        torch.cuda.current_stream(self.device).synchronize()
        return self.metrics_from_host()

    def enqueue_metrics_read(self) -> None:
        """Async copy of the accumulators into pinned host memory (current stream); pair with metrics_from_host()."""
        self._metrics_host.copy_(self._metrics, non_blocking=True)

    def metrics_from_host(self):
        """Per-client (mean batch loss, accuracy, samples) from the pinned copy; the copy must have completed."""
        h = self._metrics_host
        loss_sum = h[0].view(torch.float32).double()
        nb, ns = h[2].double().clamp(min=1), h[3].double().clamp(min=1)
        return loss_sum / nb, h[1].double() / ns, h[3].long()

    # ---- single-step entries used by tests / evaluation ---------------------------------------------------
    def forward_backward(self, learning_rate: float = 0.001, optimizer_type: str = "adam") -> None:
        self._fill_args(learning_rate, optimizer_type, train=True)
        with torch.cuda.device(self.device):
            st = L.stream_ptr(self.device)
            L.call("flb_train_begin_epoch", C.byref(self.args), st)
            L.call("flb_train_forward_backward", C.byref(self.args), st)

    def ws_array(self, name: str, dtype, per_client: int) -> torch.Tensor:
        off = L.call_ll("flb_train_ws_offset", self.model_id, self.K, self.B, name.encode())
        nbytes = self.K * self.B * per_client * torch.empty(0, dtype=dtype).element_size()
        return self.ws[off:off + nbytes].view(dtype).view(self.K, self.B, per_client)

    def evaluate(self):
        """Forward-only pass over the loaded samples (eval mode: no dropout).  Returns per-client
        (mean batch loss, accuracy, samples) and the logits [sum N, classes] on the device."""
        self._fill_args(0.0, "adam", train=False)
        logits = torch.empty((self.x.shape[0], 10), dtype=torch.float32, device=self.device)
        lg = self.ws_array("logits", torch.float32, 10)
        with torch.cuda.device(self.device):
            st = L.stream_ptr(self.device)
            ap = C.byref(self.args)
            L.call("flb_train_begin_epoch", ap, st)
            offs = self.sample_off.tolist()
            for s in range(self.max_steps()):
                L.call("flb_train_forward", ap, st)
                for k, n in enumerate(self.n_host):
                    b = min(self.B, n - s * self.B)
                    if b > 0:
                        logits[offs[k] + s * self.B: offs[k] + s * self.B + b].copy_(lg[k, :b])
                L.call("flb_train_advance", ap, st)
        loss, acc, seen = self.epoch_metrics()
        return loss.tolist(), acc.tolist(), seen.tolist(), logits


# ----------------------------------------------------------------------------------------------------------
def _collect_loader(loader: Iterable, max_batch: int = MAX_BATCH) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """Drain one pass of a DataLoader-like iterable of (data, targets) into contiguous tensors.  The kernels consume
    consecutive slices of a fixed batch size, so all batches but the last must have the same size."""
    xs, ys, sizes = [], [], []
    for data, targets in loader:
        xs.append(data)
        ys.append(targets)
        sizes.append(int(targets.shape[0]))
    if not xs:
        raise ValueError("empty data loader")
    bs = sizes[0]
    if any(s != bs for s in sizes[:-1]) or sizes[-1] > bs:
        raise ValueError(f"batches must share one size except a smaller last one, got sizes {sorted(set(sizes))}")
    if bs > max_batch:
        # The kernels map one minibatch to one 32-column MMA operand.  A larger loader batch is trained as consecutive
        # minibatches of 32 (same samples, same order, more optimizer steps) rather than refused.
        logger.warning(f"loader batch size {bs} exceeds the kernel limit {max_batch}: training in minibatches of {max_batch}")
        bs = max_batch
    return torch.cat(xs), torch.cat(ys), bs


def _engine_for(model: FederatedCNNBase, device, batch_size: int, precision: str, role: str = "train") -> BatchedClientTrainer:
    """K = 1 engine cached on the module.  ``role`` keeps evaluation apart from training: a validation pass between two
    epochs must not touch the live optimizer state (step counts feed Adam's bias correction and the Philox counters)."""
    cache = model.__dict__.setdefault("_flb_engines", {})
    key = (str(device), batch_size, precision, role)
    if key not in cache:
        cache[key] = BatchedClientTrainer(model.model_name, 1, device, batch_size, model.dropout_rate, precision)
    return cache[key]


def forward_logits(model: FederatedCNNBase, x: torch.Tensor, precision: str = "fp32") -> torch.Tensor:
    """``model(x)`` through the CUDA forward kernels.  Dropout follows ``model.training`` like nn.Dropout."""
    p = next(model.parameters())
    if not p.is_cuda:
        raise L.FlbError("model.forward: parameters are on the CPU; move the model to a CUDA device (no CPU path)")
    eng = _engine_for(model, p.device, MAX_BATCH, precision, "eval")
    _push_model(model, eng)
    eng.load_data([x.detach().to(torch.float32)], [torch.zeros(x.shape[0], dtype=torch.int64)])
    if model.training and model.dropout_rate > 0:
        raise L.FlbError("model.forward in train mode with dropout is only available inside LocalTrainer")
    return eng.evaluate()[3].to(x.device if x.is_cuda else p.device)


def _push_model(model: FederatedCNNBase, eng: BatchedClientTrainer) -> None:
    """Module parameters (+ client-local BatchNorm buffers, which are never federated) -> row 0 of the engine."""
    eng.set_client_weights(0, {n: p.data for n, p in model.named_parameters()})
    if eng.bn_running is not None:
        half = eng.bn_running.shape[1] // 2
        off = 0
        for m in model.modules():
            if isinstance(m, nn.BatchNorm2d):
                c = m.num_features
                eng.bn_running[0, off:off + c].copy_(m.running_mean)
                eng.bn_running[0, half + off:half + off + c].copy_(m.running_var)
                off += c


class LocalTrainer:
    """Drop-in for src/shared/training.py:28-403."""

    def __init__(self, model: FederatedCNNBase, device: Optional[torch.device] = None,
                 checkpoint_dir: Optional[str] = None, precision: str = "fp32"):
        self.model = model
        self.device = _cuda_device(device)
        self.checkpoint_dir = checkpoint_dir
        self.precision = precision
        L.ensure_device(self.device)            # no CPU fallback: fail here, loudly
        self.model.to(self.device)
        self.current_epoch = 0
        self.training_history: List[Dict[str, Any]] = []
        self._last_grads: Dict[str, torch.Tensor] = {}
        if self.checkpoint_dir:
            os.makedirs(self.checkpoint_dir, exist_ok=True)

    def _bn_modules(self):
        return [m for m in self.model.modules() if isinstance(m, nn.BatchNorm2d)]

    def _push(self, eng: BatchedClientTrainer) -> None:
        _push_model(self.model, eng)

    def _pull(self, eng: BatchedClientTrainer, steps: int = 0) -> None:
        views = eng.layout.views(eng.W[0])
        with torch.no_grad():
            for n, p in self.model.named_parameters():
                p.data.copy_(views[n])
            if eng.bn_running is not None:
                half = eng.bn_running.shape[1] // 2
                off = 0
                for m in self._bn_modules():
                    c = m.num_features
                    m.running_mean.copy_(eng.bn_running[0, off:off + c])
                    m.running_var.copy_(eng.bn_running[0, half + off:half + off + c])
                    m.num_batches_tracked += steps
                    off += c

    def train_local_model(self, train_loader, epochs: int, learning_rate: float = 0.001,
                          optimizer_type: str = "adam", loss_function: Optional[nn.Module] = None,
                          validation_loader=None, save_checkpoints: bool = True,
                          early_stopping_patience: Optional[int] = None) -> TrainingMetrics:
        try:
            start = time.time()
            if loss_function is not None and not isinstance(loss_function, nn.CrossEntropyLoss):
                raise ValueError("only nn.CrossEntropyLoss (the reference default) is implemented in the kernels")
            if optimizer_type.lower() not in OPTIMIZERS:
                raise ValueError(f"Unknown optimizer type: {optimizer_type}")
            self.model.train()
            eng = None
            losses, accs = [], []
            total_samples = 0
            best_val, patience = float("inf"), 0
            for epoch in range(epochs):
                self.current_epoch = epoch
                x, y, bs = _collect_loader(train_loader)          # one pass of the loader = one epoch (shuffling included)
                if eng is None:
                    eng = _engine_for(self.model, self.device, bs, self.precision)
                    eng.dropout_rate = float(self.model.dropout_rate)
                    eng.keep_last_grads = True
                    self._push(eng)
                    eng.M.zero_(); eng.V.zero_(); eng.tcount.zero_()       # fresh optimizer (training.py:89)
                elif eng.B != bs:
                    raise ValueError("batch size changed between epochs")
                eng.load_data([x], [y])
                eng._fill_args(learning_rate, optimizer_type, train=True)
                eng._run_epoch()
                loss, acc, seen = eng.epoch_metrics()
                losses.append(float(loss[0])); accs.append(float(acc[0]))
                total_samples += int(seen[0])
                self._pull(eng, steps=eng.max_steps())
                val_loss = None
                if validation_loader:
                    val_loss, _ = self._validate_epoch(validation_loader)
                    self.model.train()
                if save_checkpoints and self.checkpoint_dir:
                    self._save_checkpoint(epoch, losses[-1], val_loss)
                if early_stopping_patience and validation_loader:
                    if val_loss < best_val:
                        best_val, patience = val_loss, 0
                    else:
                        patience += 1
                        if patience >= early_stopping_patience:
                            break
            if eng is not None and eng.last_grads is not None:
                self._last_grads = {n: v.clone() for n, v in eng.layout.views(eng.last_grads[0]).items()}
            metrics = TrainingMetrics(loss=losses[-1] if losses else 0.0, accuracy=accs[-1] if accs else 0.0,
                                      epochs_completed=len(losses), training_time=time.time() - start,
                                      samples_processed=total_samples)
            self.training_history.append({"timestamp": datetime.now().isoformat(), "epochs": len(losses),
                                          "final_loss": metrics.loss, "final_accuracy": metrics.accuracy,
                                          "training_time": metrics.training_time, "samples_processed": total_samples})
            return metrics
        except Exception as e:
            logger.error(f"Local training failed: {str(e)}")
            raise TrainingError(f"Local training failed: {str(e)}")

    def _validate_epoch(self, val_loader, criterion=None) -> Tuple[float, float]:
        self.model.eval()
        x, y, bs = _collect_loader(val_loader)
        eng = _engine_for(self.model, self.device, bs, self.precision, "eval")
        self._push(eng)
        eng.load_data([x], [y])
        loss, acc, _, _ = eng.evaluate()
        return float(loss[0]), float(acc[0])

    def evaluate_model(self, test_loader) -> Dict[str, float]:
        try:
            self.model.eval()
            x, y, bs = _collect_loader(test_loader)
            eng = _engine_for(self.model, self.device, bs, self.precision, "eval")
            self._push(eng)
            eng.load_data([x], [y])
            _, _, _, logits = eng.evaluate()
            pred = logits.argmax(1).cpu()
            y = y.cpu().long()
            hit = pred == y
            out = {"overall_accuracy": hit.float().mean().item() if y.numel() else 0.0,
                   "total_samples": int(y.numel()), "correct_predictions": int(hit.sum())}
            for cls in torch.unique(y).tolist():
                m = y == cls
                out[f"class_{cls}_accuracy"] = int((hit & m).sum()) / int(m.sum())
            return out
        except Exception as e:
            raise TrainingError(f"Model evaluation failed: {str(e)}")

    def get_model_gradients(self) -> Dict[str, torch.Tensor]:
        """Gradients of the last minibatch step, as ``param.grad`` holds them upstream after ``train_local_model``
        (training.py:362-371).  The kernels keep gradients in flat rows that the optimizer consumes; the epoch's last step
        copies them out first (``flb_train_step_grads``)."""
        return {n: g.clone() for n, g in self._last_grads.items()}

    def set_model_gradients(self, gradients: Dict[str, torch.Tensor]):
        for n, p in self.model.named_parameters():
            if n in gradients:
                p.grad = gradients[n].clone().to(p.device)

    def _save_checkpoint(self, epoch: int, train_loss: float, val_loss: Optional[float] = None):
        if not self.checkpoint_dir:
            return
        ckpt = {"epoch": epoch, "model_state_dict": self.model.state_dict(), "train_loss": train_loss,
                "val_loss": val_loss, "timestamp": datetime.now().isoformat(), "model_info": self.model.get_model_info()}
        torch.save(ckpt, os.path.join(self.checkpoint_dir, f"checkpoint_epoch_{epoch}.pt"))
        torch.save(ckpt, os.path.join(self.checkpoint_dir, "latest_checkpoint.pt"))

    def load_checkpoint(self, checkpoint_path: str) -> Dict[str, Any]:
        try:
            ckpt = torch.load(checkpoint_path, map_location=self.device)
            self.model.load_state_dict(ckpt["model_state_dict"])
            self.current_epoch = ckpt["epoch"]
            return {"epoch": ckpt["epoch"], "train_loss": ckpt["train_loss"], "val_loss": ckpt.get("val_loss"),
                    "timestamp": ckpt.get("timestamp")}
        except Exception as e:
            raise TrainingError(f"Failed to load checkpoint: {str(e)}")

    def get_training_history(self) -> List[Dict[str, Any]]:
        return self.training_history.copy()

    def save_training_history(self, filepath: str):
        try:
            with open(filepath, "w") as f:
                json.dump(self.training_history, f, indent=2)
        except Exception as e:
            logger.error(f"Failed to save training history: {str(e)}")

    def reset_training_state(self):
        self.current_epoch = 0
        self.training_history = []


@dataclass
class FederatedTrainingConfig:
    """src/shared/training.py:406-452 (same keyword order and defaults)."""
    local_epochs: int = 5
    batch_size: int = 32
    learning_rate: float = 0.001
    optimizer_type: str = "adam"
    early_stopping_patience: Optional[int] = None
    save_checkpoints: bool = True
    validation_split: float = 0.1

    def to_dict(self) -> Dict[str, Any]:
        return dict(self.__dict__)

    @classmethod
    def from_dict(cls, config_dict: Dict[str, Any]) -> "FederatedTrainingConfig":
        return cls(**config_dict)


def create_adaptive_config(client_capabilities: Dict[str, Any]) -> FederatedTrainingConfig:
    """Per-client hyper-parameters from its declared capabilities (src/shared/training.py:455-501): host logic only.
    The table is upstream's, including the batch sizes 64 / 128 for 'high' power or large clients; ``LocalTrainer`` trains such a
    loader in minibatches of ``MAX_BATCH`` = 32 (the kernel limit, flb.h) and logs that it does."""
    power = client_capabilities.get("compute_power", "medium")
    bandwidth = client_capabilities.get("network_bandwidth", 10)
    samples = client_capabilities.get("available_samples", 1000)
    epochs, batch, lr = {"high": (10, 64, 0.001), "medium": (5, 32, 0.001)}.get(power, (3, 16, 0.0005))
    if samples < 500:
        batch = min(batch, 16)
    elif samples > 5000:
        batch = min(batch * 2, 128)
    if bandwidth < 5:                       # slow uplink: more local work per round
        epochs = max(epochs + 2, 7)
    return FederatedTrainingConfig(local_epochs=epochs, batch_size=batch, learning_rate=lr, optimizer_type="adam",
                                   early_stopping_patience=None, save_checkpoints=True, validation_split=0.1)


def validate_training_data(train_loader) -> Dict[str, Any]:
    """Pre-flight check of a loader (src/shared/training.py:504-560): never raises, returns {'valid': ...}."""
    try:
        if len(train_loader) == 0:
            raise ValueError("Training data loader is empty")
        batch = next(iter(train_loader))
        if len(batch) != 2:
            raise ValueError("Expected (data, targets) tuple from data loader")
        data, targets = batch
        if not isinstance(data, torch.Tensor):
            raise ValueError("Data must be a torch.Tensor")
        if not isinstance(targets, torch.Tensor):
            raise ValueError("Targets must be a torch.Tensor")
        if data.dim() != 4:
            raise ValueError(f"Expected 4D data tensor, got shape {data.shape}")
        return {"valid": True, "num_batches": len(train_loader), "batch_size": data.shape[0], "data_shape": tuple(data.shape[1:]),
                "num_classes": len(torch.unique(targets)), "data_type": str(data.dtype), "targets_type": str(targets.dtype)}
    except Exception as e:
        logger.error(f"Training data validation failed: {str(e)}")
        return {"valid": False, "error": str(e)}
