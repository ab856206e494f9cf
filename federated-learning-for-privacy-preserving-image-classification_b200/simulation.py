"""The federated round, in process and batched: what the reference's simulation does with N client threads, an
in-process gRPC server and 1 s / 10 s sleeps (src/simulation/federated_simulation.py:194-527), restated as

    for every resident client at once:   w_k <- global                       (federated_trainer.py:367-388)
                                         w_k <- LocalTrainer(w_k, data_k)    (federated_trainer.py:390-426)
                                         u_k <- global + DP(w_k - global)    (federated_trainer.py:428-469)
    global <- FedAvg({u_k}, n_k)                                             (fedavg.py:56-124)

Clients are rows of device matrices; client i lives on rank i mod world_size; each rank aggregates its own rows
with globally normalised weights n_k / sum(n) and the client -> coordinator hop is ONE NCCL all-reduce of P floats
(which doubles as the next round's broadcast).  ``SimulationConfig`` / ``FederatedLearningSimulation`` /
``run_mnist_simulation`` keep the reference's signatures and result-dict keys (:32-42, :362-405, :430-469, :530-583)."""
from __future__ import annotations

import logging
import os
import time
from datetime import datetime
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from . import ops
from .models_pytorch import INPUT_SHAPES, ModelFactory
from .privacy import PrivacyError, create_privacy_engine
from .training import BatchedClientTrainer

logger = logging.getLogger(__name__)

MNIST_SIZES = (480, 512, 544, 576)      # non-uniform so that FedAvg weights are non-trivial
CIFAR_SIZES = (416, 448, 480)


def synthetic_num_samples(model_name: str, client_idx: int) -> int:
    sizes = MNIST_SIZES if model_name == "simple_cnn" else CIFAR_SIZES
    return sizes[client_idx % len(sizes)]


def synthetic_client_data(model_name: str, client_idx: int, n: Optional[int] = None):
    """x ~ N(0,1) fp32 [N_c, C, H, W], y ~ U{0..9} int64, generator seed 1000 + client_idx."""
    g = torch.Generator().manual_seed(1000 + client_idx)
    n = n or synthetic_num_samples(model_name, client_idx)
    x = torch.randn((n,) + INPUT_SHAPES[model_name], generator=g, dtype=torch.float32)
    y = torch.randint(0, 10, (n,), generator=g, dtype=torch.int64)
    return x, y


def shard_clients(num_clients: int, rank: int, world_size: int) -> List[int]:
    """client i -> rank i mod G."""
    return list(range(rank, num_clients, world_size))


def global_fedavg_weights(num_samples_all: Sequence[int], client_ids: Sequence[int]) -> List[float]:
    """n_k / sum over ALL clients (fedavg.py:247-256): each rank's partial sum is already globally normalised, so the
    cross-GPU step is a plain sum."""
    total = sum(num_samples_all)
    if total == 0:
        return [1.0 / len(num_samples_all)] * len(client_ids)
    return [num_samples_all[i] / total for i in client_ids]


class FederatedRoundEngine:
    """All of one rank's clients for one model: train -> update-level DP -> FedAvg (-> all-reduce)."""

    def __init__(self, model_name: str, num_clients: int, device=None, rank: int = 0, world_size: int = 1,
                 process_group=None, batch_size: int = 32, local_epochs: int = 1, learning_rate: float = 0.001,
                 optimizer_type: str = "adam", dp_mode: str = "update", epsilon: float = 1.0, delta: float = 1e-5,
                 max_grad_norm: float = 1.0, dropout_rate: Optional[float] = None, precision: str = "fp32",
                 compression: Optional[str] = None, seed: Optional[int] = None, use_graph: bool = True,
                 topk_sparsity: float = 0.9):
        if dp_mode not in ("none", "update", "per_sample"):
            raise ValueError("dp_mode must be 'none', 'update' (reference behaviour) or 'per_sample'")
        if compression not in (None, "q8", "topk"):
            raise ValueError("compression must be None, 'q8' (QuantizationCompressor, 8 bit) or 'topk' (TopKSparsificationCompressor)")
        if not (0.0 <= topk_sparsity <= 1.0):
            raise ValueError("topk_sparsity must be in [0, 1]")
        self.topk_sparsity = float(topk_sparsity)
        self._topk_meta = None
        self.model_name = model_name
        self.num_clients = int(num_clients)
        self.rank, self.world_size, self.pg = rank, world_size, process_group
        self.client_ids = shard_clients(self.num_clients, rank, world_size)
        if not self.client_ids:
            raise ValueError(f"rank {rank} owns no client ({num_clients} clients over {world_size} ranks)")
        self.local_epochs, self.lr, self.optimizer_type = local_epochs, learning_rate, optimizer_type
        self.dp_mode, self.epsilon, self.delta, self.max_grad_norm = dp_mode, epsilon, delta, max_grad_norm
        self.compression = compression
        # One Philox seed for the round's DP noise and the trainer's dropout masks.  None = fresh OS entropy (rank 0's
        # draw is shared across ranks so that the streams, keyed by GLOBAL client index, do not depend on the GPU count).
        self.seed = self._agree_on_seed(seed)
        seed = self.seed
        self.trainer = BatchedClientTrainer(model_name, len(self.client_ids), device, batch_size, dropout_rate,
                                            precision, seed=seed, client_base=rank, client_stride=world_size, use_graph=use_graph)
        self.device = self.trainer.device
        self.layout = self.trainer.layout
        if dp_mode == "per_sample":
            self.trainer.configure_dp("per_sample", max_grad_norm,
                                      max_grad_norm * ops.gaussian_sigma_unit(epsilon, delta))
        # Several ranks: the aggregate is formed by ONE kernel per rank over NVLink peer memory (p2p.PeerFedAvg: FedAvg partial
        # sum fused with the cross-GPU reduction), and the global row lives in the peer-mapped region.  FLB_NO_P2P=1, or peer
        # memory being unavailable (all ranks agree on that inside the constructor), selects FedAvg kernel + NCCL all_reduce.
        self.p2p = None
        if world_size > 1 and not os.environ.get("FLB_NO_P2P"):
            from .p2p import PeerFedAvg
            try:
                self.p2p = PeerFedAvg(self.layout.ld, self.device, rank, world_size, process_group)
            except L.FlbError as e:
                logging.getLogger(__name__).info("peer-memory FedAvg unavailable (%s): FedAvg kernel + NCCL all_reduce", e)
                self.p2p = None
        self.global_row = (self.p2p.global_row if self.p2p is not None
                           else torch.zeros(self.layout.ld, dtype=torch.float32, device=self.device))
        self._stage_row = torch.zeros(self.layout.ld, dtype=torch.float32, device=self.device) if self.p2p is not None else None
        self.upload = self.layout.new_rows(len(self.client_ids), self.device)
        self.norms: Optional[torch.Tensor] = None
        self.num_samples_all: List[int] = []
        self.round_number = 0
        self.history: List[Dict[str, Any]] = []
        self.dp_z: Optional[torch.Tensor] = None             # injected standard normals [K_local, ld] (parity tests)

    def _agree_on_seed(self, seed: Optional[int]) -> int:
        if seed is not None:
            return int(seed) & (2**64 - 1)
        seed = L.fresh_seed()
        if self.world_size > 1:
            import torch.distributed as dist
            box = [seed]
            dist.broadcast_object_list(box, src=0, group=self.pg)
            seed = int(box[0])
        return seed

    # ---- setup -------------------------------------------------------------------------------------------
    def set_global_weights(self, weights: Dict[str, torch.Tensor]) -> None:
        self.global_row.zero_()
        self.layout.flatten_into(self.global_row, weights)

    def global_weights(self, device=None) -> Dict[str, torch.Tensor]:
        return self.layout.unflatten(self.global_row, device)

    def load_synthetic(self, sizes: Optional[Sequence[int]] = None) -> None:
        xs, ys = [], []
        for i in self.client_ids:
            x, y = synthetic_client_data(self.model_name, i, None if sizes is None else sizes[i])
            xs.append(x)
            ys.append(y)
        all_sizes = [sizes[i] if sizes is not None else synthetic_num_samples(self.model_name, i)
                     for i in range(self.num_clients)]
        self.load_data(xs, ys, all_sizes)

    def load_data(self, xs, ys, num_samples_all: Sequence[int]) -> None:
        """xs / ys: this rank's clients (in ``client_ids`` order); num_samples_all: sample counts of ALL clients, so
        that every rank can form the global FedAvg weights without a collective."""
        self.trainer.load_data(xs, ys)
        self.num_samples_all = [int(n) for n in num_samples_all]

    def load_packed(self, x_all: torch.Tensor, y_all: torch.Tensor, num_samples_all: Sequence[int]) -> None:
        """This rank's clients concatenated (pinned host memory recommended): two H2D copies per round."""
        self.trainer.load_packed(x_all, y_all, [int(num_samples_all[i]) for i in self.client_ids])
        self.num_samples_all = [int(n) for n in num_samples_all]

    def attach_device_shards(self, x_dev: torch.Tensor, y_dev: torch.Tensor, num_samples_all: Sequence[int]) -> None:
        """Device-resident packed store built by ``data_loader.DeviceShardBuilder`` for this rank's clients: used in place."""
        self.trainer.attach(x_dev, y_dev, [int(num_samples_all[i]) for i in self.client_ids])
        self.num_samples_all = [int(n) for n in num_samples_all]

    def prefetch_packed(self, x_all: torch.Tensor, y_all: torch.Tensor) -> None:
        """Start uploading the next round's samples (same client sizes) on the copy stream; see BatchedClientTrainer."""
        self.trainer.prefetch_packed(x_all, y_all, self.trainer.n_host)

    def use_prefetched(self) -> None:
        self.trainer.use_prefetched()

    # ---- one round -----------------------------------------------------------------------------------------
    def fedavg_weights(self) -> List[float]:
        """n_k * E / sum(n * E) over ALL clients (fedavg.py:247-256 with num_samples = samples_processed)."""
        return global_fedavg_weights(self.num_samples_all, self.client_ids)

    def run_round(self, read_metrics: bool = True, model_out: Optional[torch.Tensor] = None) -> Dict[str, Any]:
        """One FedAvg round.  ``model_out``: optional pinned host buffer [>= P] that receives the aggregated model (the
        copy is enqueued before the round's single synchronisation, so metrics and model arrive together)."""
        self.start_round()
        return self.finish_round(read_metrics, model_out)

    def start_round(self) -> None:
        """Enqueue the whole round (local epochs, DP, FedAvg, all-reduce) on the current stream; does not synchronise.
        Host work that should overlap the round goes between ``start_round`` and ``finish_round``.  (Issue
        ``prefetch_packed`` BEFORE ``start_round``: a copy submitted behind an already enqueued captured epoch only starts
        when that epoch has drained -- measured with scripts/dbg_e2e.py.)"""
        tr, lay = self.trainer, self.layout
        self._t0 = time.time()
        tr.begin_round(self.global_row)                    # global model into every client row, fresh optimizer state: one launch
        tr._fill_args(self.lr, self.optimizer_type, train=True)
        for _ in range(self.local_epochs):
            tr._run_epoch()
        rows = tr.W
        absmax = None
        if self.dp_mode == "update":
            # fresh engine per client-round upstream (budget semantics); Philox stream = global
            # client index + round * num_clients so results do not depend on how clients are spread over GPUs.
            # With the uint8 codec the clip + noise pass also reduces the quantiser's per-layer max|upload|.
            sigma_unit = ops.gaussian_sigma_unit(self.epsilon, self.delta)
            res = ops.dp_clip_noise(tr.W, self.global_row, self.max_grad_norm, sigma_unit,
                                    seed=self.seed ^ 0x0DD5EED,
                                    stream_base=self.round_number * self.num_clients + self.rank,
                                    stream_stride=self.world_size, z=self.dp_z, P=lay.P, out=self.upload,
                                    absmax_seg=lay.seg_off(self.device) if self.compression == "q8" and len(lay.names) <= 64 else None)
            rows, self.norms = res[0], res[1]
            absmax = res[2] if len(res) > 2 else None
        w = self.fedavg_weights()
        if self.p2p is not None and self.compression is None:
            self.p2p.reduce(rows, w, lay.P)               # partial sum + cross-GPU reduction + broadcast: one kernel
            self.round_number += 1
            self._round_w = w
            return
        # the aggregate is written straight into the global row (every client row was copied from it at the start of the
        # round; nothing reads it until the next one); with peer memory the codec paths stage their partial sum first
        partial = self.global_row[:lay.P] if self.p2p is None else self._stage_row[:lay.P]
        if self.compression == "q8":
            seg = lay.seg_off(self.device)
            q, scale, zp = ops.q8_quantize(rows, seg, P=lay.P, absmax=absmax)
            partial = ops.fedavg_weighted_sum_q8(q, scale, zp, seg, w, lay.P, out=partial)
        elif self.compression == "topk":
            # TopKSparsificationCompressor (compression.py:327-365) on every upload: keep the k = max(1, int(n * (1 - sparsity)))
            # largest-magnitude entries of each layer, zeros elsewhere, then FedAvg over the reconstructed rows
            seg = lay.seg_off(self.device)
            if self._topk_meta is None:
                import math
                kk = [ops.topk_count(math.prod(lay.shapes[n]), self.topk_sparsity) for n in lay.names]
                self._topk_meta = kk
            idx, val, off_t, kk_t = ops.topk_select(rows, seg, self._topk_meta, P=lay.P)
            dense = ops.topk_scatter(idx, val, seg, kk_t, off_t, lay.P, lay.ld)
            ops.fedavg_weighted_sum(dense, w, P=lay.P, out=partial)
        else:
            ops.fedavg_weighted_sum(rows, w, P=lay.P, out=partial)
        if self.p2p is not None:
            self.p2p.reduce(self._stage_row, [1.0], lay.P)                # 0 + 1 * x is exact: the same kernel, one row
        elif self.world_size > 1:
            # the rank's partial sum lands directly in the global row and is all-reduced in place: the collective is the
            # client -> coordinator hop AND the next round's broadcast; no staging copy on either side of it
            import torch.distributed as dist
            dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=self.pg)
        self.round_number += 1
        self._round_w = w

    def finish_round(self, read_metrics: bool = True, model_out: Optional[torch.Tensor] = None) -> Dict[str, Any]:
        tr, lay = self.trainer, self.layout
        out: Dict[str, Any] = {"round": self.round_number, "clients": len(self.client_ids)}
        if model_out is not None:
            model_out[:lay.P].copy_(self.global_row[:lay.P], non_blocking=True)
        if read_metrics:
            loss, acc, seen = tr.epoch_metrics()          # the round's single synchronisation (covers model_out too)
            w = self._round_w
            out.update(losses=loss.tolist(), accuracies=acc.tolist(),
                       samples=[int(s) * self.local_epochs for s in seen.tolist()],
                       avg_loss=float(sum(l * wi for l, wi in zip(loss.tolist(), w))), wall_s=time.time() - self._t0)
            self.history.append(out)
        return out

    def evaluate_global(self, x: torch.Tensor, y: torch.Tensor, shards: int = 8) -> Dict[str, float]:
        """Accuracy / mean batch loss of the CURRENT global model on a held-out set, through the forward kernels in eval
        mode (no dropout, BatchNorm running statistics) -- what ``LocalTrainer.evaluate_model`` / ``_validate_epoch`` compute
        (src/shared/training.py:214-242, 307-360), batched: the set is cut into ``shards`` pseudo-clients that all carry the
        global weights, so one launch sequence evaluates ``shards x batch`` samples.  Uses its own small trainer; the
        round state is untouched.  This is the number ``GlobalModel.accuracy_metrics`` is meant to hold (fedavg.py:106)."""
        n = int(x.shape[0])
        shards = max(1, min(shards, (n + self.trainer.B - 1) // self.trainer.B))
        ev = getattr(self, "_evaluator", None)
        if ev is None or ev.K != shards:
            ev = BatchedClientTrainer(self.model_name, shards, self.device, self.trainer.B, 0.0, self.trainer.precision)
            self._evaluator = ev
        ev.set_global_row(self.global_row)
        if ev.bn_running is not None:               # BatchNorm buffers are client-local upstream: use client 0's as the server's
            ev.bn_running.copy_(self.trainer.bn_running[:1].expand(shards, -1))
        per = (n + shards - 1) // shards
        xs = [x[i * per:(i + 1) * per] for i in range(shards)]
        ys = [y[i * per:(i + 1) * per] for i in range(shards)]
        ev.load_data(xs, ys)
        loss, acc, seen, _ = ev.evaluate()
        tot = max(sum(seen), 1)
        return {"accuracy": float(sum(a * s for a, s in zip(acc, seen)) / tot),
                "loss": float(sum(l * s for l, s in zip(loss, seen)) / tot), "samples": int(sum(seen))}

    def samples_per_round(self) -> int:
        return sum(self.trainer.n_host) * self.local_epochs


# ----------------------------------------------------------------------------------------------------------
class SimulationConfig:
    """Same constructor as src/simulation/federated_simulation.py:32-42; host/port are accepted and unused."""

    def __init__(self, num_clients: int = 5, num_rounds: int = 10, dataset_name: str = "mnist",
                 model_type: str = "simple_cnn", partition_strategy: str = "non_iid", target_accuracy: float = 0.91,
                 coordinator_host: str = "localhost", coordinator_port: int = 50051, privacy_epsilon: float = 1.0,
                 privacy_delta: float = 1e-5):
        self.num_clients = num_clients
        self.num_rounds = num_rounds
        self.dataset_name = dataset_name
        self.model_type = model_type
        self.partition_strategy = partition_strategy
        self.target_accuracy = target_accuracy
        self.coordinator_host = coordinator_host
        self.coordinator_port = coordinator_port
        self.privacy_epsilon = privacy_epsilon
        self.privacy_delta = privacy_delta

    def to_dict(self) -> Dict[str, Any]:
        return dict(self.__dict__)


class FederatedLearningSimulation:
    """``run_simulation`` returns the reference's result dict (keys simulation_config, start_time, end_time,
    duration_seconds, success, clients, summary -- :430-469) or ``{'error': ...}``; it never raises.  Data is the
    seeded synthetic set of SURVEY.md 8(d) (no dataset download exists offline); extra keyword arguments select the
    engine options and default to the reference's behaviour (update-level DP, Adam 1e-3, 5 local epochs, batch 32)."""

    def __init__(self, config: SimulationConfig, device=None, local_epochs: int = 5, dp_mode: str = "update",
                 precision: str = "fp32", data: Optional[Tuple[list, list]] = None, **engine_kw):
        self.config = config
        self.device = device
        self.local_epochs = local_epochs
        self.dp_mode = dp_mode
        self.precision = precision
        self.data = data
        self.engine_kw = engine_kw
        self.engine: Optional[FederatedRoundEngine] = None
        self.simulation_results: Dict[str, Any] = {}
        self.start_time = self.end_time = None

    def run_simulation(self, timeout_minutes: int = 60) -> Dict[str, Any]:
        cfg = self.config
        try:
            self.start_time = datetime.now()
            eng = FederatedRoundEngine(cfg.model_type, cfg.num_clients, self.device, local_epochs=self.local_epochs,
                                       dp_mode=self.dp_mode, epsilon=cfg.privacy_epsilon, delta=cfg.privacy_delta,
                                       precision=self.precision, **self.engine_kw)
            self.engine = eng
            torch.manual_seed(0)
            eng.set_global_weights(ModelFactory.create_model(cfg.model_type).get_model_weights())
            if self.data is not None:
                eng.load_data(self.data[0], self.data[1], [int(x.shape[0]) for x in self.data[0]])
            else:
                eng.load_synthetic()
            deadline = time.time() + 60 * timeout_minutes
            rounds = []
            for _ in range(cfg.num_rounds):
                rounds.append(eng.run_round())
                if time.time() > deadline:
                    break
            self.end_time = datetime.now()
            clients = {}
            for j, cid in enumerate(eng.client_ids):
                clients[f"client_{cid}"] = {
                    "status": {"is_running": False, "rounds_completed": len(rounds), "client_id": f"client_{cid}"},
                    "training_history": [{"round": r["round"], "final_loss": r["losses"][j],
                                          "final_accuracy": r["accuracies"][j],
                                          "samples_processed": r["samples"][j]} for r in rounds]}
            acc = float(np.average(rounds[-1]["accuracies"], weights=eng.fedavg_weights())) if rounds else 0.0
            best = max((float(np.average(r["accuracies"], weights=eng.fedavg_weights())) for r in rounds), default=0.0)
            summary = {"final_accuracy": acc, "best_accuracy": best, "total_rounds": len(rounds),
                       "avg_round_duration": float(np.mean([r["wall_s"] for r in rounds])) if rounds else 0.0,
                       "target_achieved": acc >= cfg.target_accuracy, "total_clients": len(clients), "active_clients": 0,
                       "avg_participation_rate": 1.0, "min_participation_rate": 1.0, "max_participation_rate": 1.0,
                       "privacy_preserved": True, "privacy_epsilon": cfg.privacy_epsilon,
                       "privacy_delta": cfg.privacy_delta}
            self.simulation_results = {
                "simulation_config": cfg.to_dict(), "start_time": self.start_time.isoformat(),
                "end_time": self.end_time.isoformat(),
                "duration_seconds": (self.end_time - self.start_time).total_seconds(), "success": True,
                "training_progress": [{"round": r["round"], "avg_loss": r["avg_loss"]} for r in rounds],
                "clients": clients, "summary": summary}
            return self.simulation_results
        except Exception as e:
            logger.error(f"Simulation failed: {e}")
            return {"error": str(e)}


def run_mnist_simulation(num_clients: int = 5, num_rounds: int = 10, target_accuracy: float = 0.91, **kw) -> Dict[str, Any]:
    cfg = SimulationConfig(num_clients=num_clients, num_rounds=num_rounds, dataset_name="mnist",
                           model_type="simple_cnn", target_accuracy=target_accuracy, privacy_epsilon=1.0,
                           privacy_delta=1e-5)
    return FederatedLearningSimulation(cfg, **kw).run_simulation(timeout_minutes=30)


def run_cifar10_simulation(num_clients: int = 5, num_rounds: int = 20, target_accuracy: float = 0.75, **kw) -> Dict[str, Any]:
    cfg = SimulationConfig(num_clients=num_clients, num_rounds=num_rounds, dataset_name="cifar10",
                           model_type="cifar10_cnn", target_accuracy=target_accuracy, privacy_epsilon=1.0,
                           privacy_delta=1e-5)
    return FederatedLearningSimulation(cfg, **kw).run_simulation(timeout_minutes=60)
