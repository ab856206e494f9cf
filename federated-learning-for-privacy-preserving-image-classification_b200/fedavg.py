"""FedAvg aggregation behind the reference's ``FedAvgAggregator`` surface
(``src/aggregation/fedavg.py``: aggregate_updates :56-124, filtering :209-245, sample weights :247-256,
normalisation :258-265, the hot loop ``_weighted_average`` :267-289, stats :291-357,
``AdaptiveFedAvg`` :360-467, factory :470-484).

Host logic (filtering, weights in Python float64, history, exception types/messages) follows the
reference; the K x L axpy loop is one launch of ``csrc/fedavg.cu`` over the stacked ``[K, ld]`` client
rows.  Updates whose tensors are already on the device are read in place through a pointer table
(no gather copy); host-resident updates are staged through one pinned buffer and a single H2D copy,
and the result is returned on the device of the first update's tensors, as new tensors, like
``zeros_like(first)`` does upstream."""
from __future__ import annotations

import logging
import pickle
from collections import OrderedDict
from datetime import datetime
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from . import _lib as L
from . import ops
from .interfaces import AggregationServiceInterface
from .layout import ParamLayout
from .models import GlobalModel, ModelUpdate, ModelWeights
from .validation import ModelUpdateValidator, validate_model_compatibility

logger = logging.getLogger(__name__)


class FedAvgError(Exception):
    pass


class _Staging:
    """Pinned host + device staging for host-resident updates, reused across rounds."""

    def __init__(self):
        self.key = None
        self.host = None
        self.dev = None

    def get(self, K: int, layout: ParamLayout, device: torch.device):
        key = (K, layout.signature(), str(device))
        if key != self.key:
            self.host = torch.zeros((K, layout.ld), dtype=torch.float32).pin_memory()
            self.dev = torch.empty((K, layout.ld), dtype=torch.float32, device=device)
            self.key = key
        return self.host, self.dev


class FedAvgAggregator(AggregationServiceInterface):
    def __init__(self, min_clients: int = 2, max_clients: Optional[int] = None, validate_updates: bool = True,
                 device: Optional[torch.device] = None):
        self.min_clients = min_clients
        self.max_clients = max_clients
        self.validate_updates = validate_updates
        self.validator = ModelUpdateValidator() if validate_updates else None
        self.aggregation_history: List[Dict[str, Any]] = []
        self.device = device            # None -> current CUDA device at call time
        self._staging = _Staging()

    # ---- public surface ------------------------------------------------------------------------
    def aggregate_updates(self, updates: List[ModelUpdate], weights: Optional[List[float]] = None) -> GlobalModel:
        try:
            t0 = datetime.now()
            self._validate_aggregation_inputs(updates, weights)
            valid = self._filter_and_validate_updates(updates)
            if len(valid) < self.min_clients:
                raise FedAvgError(f"Insufficient valid updates: {len(valid)} < {self.min_clients}")
            if self.max_clients and len(valid) > self.max_clients:
                valid = sorted(valid, key=lambda u: u.num_samples, reverse=True)[:self.max_clients]
            if weights is None:
                agg_w = self._calculate_sample_weights(valid)
            else:
                agg_w = self._normalize_weights(weights[:len(valid)])        # positional slice, as upstream :92
            averaged = self._weighted_average(valid, agg_w)
            total_samples = sum(u.num_samples for u in valid)
            avg_loss = sum(u.training_loss * w for u, w in zip(valid, agg_w))
            model = GlobalModel(round_number=valid[0].round_number, model_weights=averaged, accuracy_metrics={},
                                participating_clients=[u.client_id for u in valid], convergence_score=0.0,
                                created_at=datetime.now())
            self._record_aggregation_stats(valid, agg_w, total_samples, avg_loss,
                                           (datetime.now() - t0).total_seconds())
            return model
        except Exception as e:
            logger.error(f"FedAvg aggregation failed: {str(e)}")
            raise FedAvgError(f"FedAvg aggregation failed: {str(e)}")

    def validate_update(self, update: ModelUpdate) -> bool:
        try:
            if not self.validate_updates or not self.validator:
                return True
            return self.validator.validate_model_update(update)
        except Exception as e:
            logger.error(f"Update validation failed for client {update.client_id}: {str(e)}")
            return False

    def compress_global_model(self, model: GlobalModel) -> bytes:
        return pickle.dumps(model.model_weights)

    def calculate_convergence_metrics(self, old_model: GlobalModel, new_model: GlobalModel) -> float:
        """sum_l ||new_l - old_l|| / sum_l ||new_l||, clamped to [0, 1] (fedavg.py:144-190)."""
        try:
            if not old_model or not new_model:
                return 1.0
            diff = norm = 0.0
            names = [n for n in new_model.model_weights if n in old_model.model_weights]
            nw = [new_model.model_weights[n] for n in names]
            ow = [old_model.model_weights[n] for n in names]
            if names and all(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.device == nw[0].device
                             and t.shape == o.shape for t, o in zip(nw + ow, ow + nw)):
                dev = nw[0].device                      # one launch + one read instead of two norms and syncs per layer
                offs = [0]
                for t in nw:
                    offs.append(offs[-1] + t.numel())
                sums = ops.delta_norms(torch.tensor([t.data_ptr() for t in nw], dtype=torch.int64).to(dev),
                                       torch.tensor([t.data_ptr() for t in ow], dtype=torch.int64).to(dev),
                                       torch.tensor(offs, dtype=torch.int64, device=dev), offs[-1], dev).sqrt().sum(0).cpu().tolist()
                return min(1.0, max(0.0, sums[0] / sums[1] if sums[1] > 0 else 0.0))
            for name, new_w in new_model.model_weights.items():
                if name in old_model.model_weights:
                    diff += torch.norm(new_w - old_model.model_weights[name].to(new_w.device)).item()
                    norm += torch.norm(new_w).item()
            return min(1.0, max(0.0, diff / norm if norm > 0 else 0.0))
        except Exception as e:
            logger.error(f"Convergence calculation failed: {str(e)}")
            return 0.0

    def get_aggregation_stats(self) -> Dict[str, Any]:
        if not self.aggregation_history:
            return {"message": "No aggregation history available"}
        recent = self.aggregation_history[-10:]
        return {"total_aggregations": len(self.aggregation_history), "recent_aggregations": len(recent),
                "avg_clients_per_round": np.mean([s["num_clients"] for s in recent]),
                "avg_samples_per_round": np.mean([s["total_samples"] for s in recent]),
                "avg_aggregation_time": np.mean([s["aggregation_time"] for s in recent]),
                "avg_training_loss": np.mean([s["avg_training_loss"] for s in recent]),
                "client_participation": self._calculate_client_participation()}

    # ---- host logic ----------------------------------------------------------------------------
    def _validate_aggregation_inputs(self, updates, weights):
        if not updates:
            raise FedAvgError("No model updates provided")
        if weights is not None:
            if len(weights) != len(updates):
                raise FedAvgError("Number of weights must match number of updates")
            if any(w < 0 for w in weights):
                raise FedAvgError("All weights must be non-negative")
            if sum(weights) == 0:
                raise FedAvgError("Sum of weights cannot be zero")

    def _filter_and_validate_updates(self, updates: List[ModelUpdate]) -> List[ModelUpdate]:
        valid = []
        for u in updates:
            try:
                if u.num_samples <= 0 or u.training_loss < 0:
                    logger.warning(f"Skipping update from {u.client_id}: invalid sample count or loss")
                    continue
                if self.validate_updates and not self.validate_update(u):
                    logger.warning(f"Skipping update from {u.client_id}: validation failed")
                    continue
                valid.append(u)
            except Exception as e:
                logger.error(f"Error validating update from {u.client_id}: {str(e)}")
        if len(valid) > 1:
            ref = valid[0].model_weights
            # upstream pops from the list it is enumerating (:237-243); same quirk kept: indices refer to
            # the list as it was when the scan started
            for i, u in enumerate(valid[1:], 1):
                try:
                    validate_model_compatibility(ref, u.model_weights)
                except Exception as e:
                    logger.warning(f"Removing incompatible update from {u.client_id}: {str(e)}")
                    valid.pop(i)
        return valid

    def _calculate_sample_weights(self, updates: List[ModelUpdate]) -> List[float]:
        total = sum(u.num_samples for u in updates)
        if total == 0:
            return [1.0 / len(updates)] * len(updates)
        return [u.num_samples / total for u in updates]

    def _normalize_weights(self, weights: List[float]) -> List[float]:
        total = sum(weights)
        if total == 0:
            return [1.0 / len(weights)] * len(weights)
        return [w / total for w in weights]

    # ---- the hot path ----------------------------------------------------------------------------
    def _weighted_average(self, updates: List[ModelUpdate], weights: List[float]) -> ModelWeights:
        if not updates:
            raise FedAvgError("No updates to aggregate")
        first = updates[0].model_weights
        layout = ParamLayout.from_weights(first)
        K = len(updates)
        out_device = next(iter(first.values())).device
        device = torch.device(self.device) if self.device is not None else (
            out_device if out_device.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
            if torch.cuda.is_available() else torch.device("cuda"))
        L.ensure_device(device)
        for u in updates:
            for name in u.model_weights:
                if name not in layout.offsets:
                    logger.warning(f"Layer {name} not found in reference model")
            # The kernels address every update with the FIRST update's layer sizes: a tensor of another shape (or a missing
            # one) would be read out of bounds.  Upstream fails on the shape mismatch inside `+=` (fedavg.py:285); the
            # compatibility filter can let such an update through (pop-while-enumerating, :237-243) and
            # validate_updates=False skips it altogether, so the check lives here, before any pointer is taken.
            for name in layout.names:
                t = u.model_weights.get(name)
                if t is None:
                    raise FedAvgError(f"Update from {u.client_id} has no tensor for layer {name}")
                if tuple(t.shape) != tuple(layout.shapes[name]):
                    raise FedAvgError(f"Update from {u.client_id}: layer {name} has shape {tuple(t.shape)}, "
                                      f"expected {tuple(layout.shapes[name])}")
        on_device = all(u.model_weights[n].device == device and u.model_weights[n].dtype == torch.float32
                        and u.model_weights[n].is_contiguous() for u in updates for n in layout.names)
        if on_device:
            table = torch.tensor([u.model_weights[n].data_ptr() for u in updates for n in layout.names],
                                 dtype=torch.int64).to(device, non_blocking=True)
            flat = ops.fedavg_weighted_sum_ptrs(table, layout.seg_off(device), weights, K, len(layout.names),
                                                layout.P, device)
        else:
            host, dev = self._staging.get(K, layout, device)
            dst, src = [], []
            for k, u in enumerate(updates):
                v = layout.views(host[k])
                for n in layout.names:
                    dst.append(v[n])
                    src.append(u.model_weights[n].detach().to(torch.float32))
            if all(s.device.type == "cpu" for s in src):
                torch._foreach_copy_(dst, src)
                dev.copy_(host, non_blocking=True)
            else:                                   # mixed devices: copy row by row
                for k, u in enumerate(updates):
                    layout.flatten_into(dev[k], u.model_weights)
            flat = ops.fedavg_weighted_sum(dev, weights, P=layout.P)
        return layout.unflatten(flat, out_device)

    def aggregate_rows(self, theta: torch.Tensor, weights, P: int, out: Optional[torch.Tensor] = None,
                       accumulate: bool = False) -> torch.Tensor:
        """Flat entry for the batched round driver: theta [K, ld] device rows -> [P] global row."""
        return ops.fedavg_weighted_sum(theta, weights, P=P, out=out, accumulate=accumulate)

    # ---- stats -----------------------------------------------------------------------------------
    def _record_aggregation_stats(self, updates, weights, total_samples, avg_training_loss, aggregation_time):
        self.aggregation_history.append({
            "timestamp": datetime.now().isoformat(), "num_clients": len(updates), "total_samples": total_samples,
            "avg_training_loss": avg_training_loss, "aggregation_time": aggregation_time,
            "client_weights": {u.client_id: w for u, w in zip(updates, weights)},
            "client_samples": {u.client_id: u.num_samples for u in updates}})
        if len(self.aggregation_history) > 100:
            self.aggregation_history = self.aggregation_history[-100:]

    def _calculate_client_participation(self) -> Dict[str, Any]:
        if not self.aggregation_history:
            return {}
        counts: Dict[str, int] = {}
        for s in self.aggregation_history:
            for cid in s["client_weights"]:
                counts[cid] = counts.get(cid, 0) + 1
        rounds = len(self.aggregation_history)
        return {"unique_clients": len(counts),
                "avg_participation_rate": np.mean(list(counts.values())) / rounds,
                "most_active_clients": sorted(counts.items(), key=lambda x: x[1], reverse=True)[:5]}


class AdaptiveFedAvg(FedAvgAggregator):
    """Sample weights blended with a loss-history term on the host (fedavg.py:360-467); the kernel is the same."""

    def __init__(self, min_clients: int = 2, max_clients: Optional[int] = None, validate_updates: bool = True,
                 performance_weight: float = 0.1, device: Optional[torch.device] = None):
        super().__init__(min_clients, max_clients, validate_updates, device)
        self.performance_weight = performance_weight
        self.client_performance_history: Dict[str, Dict[str, Any]] = {}

    def aggregate_updates(self, updates, weights=None) -> GlobalModel:
        try:
            self._update_performance_history(updates)
            if weights is None:
                weights = self._calculate_adaptive_weights(updates)
            return super().aggregate_updates(updates, weights)
        except Exception as e:
            raise FedAvgError(f"Adaptive FedAvg aggregation failed: {str(e)}")

    def _update_performance_history(self, updates):
        for u in updates:
            h = self.client_performance_history.setdefault(
                u.client_id, {"losses": [], "sample_counts": [], "participation_count": 0})
            h["losses"].append(u.training_loss)
            h["sample_counts"].append(u.num_samples)
            h["participation_count"] += 1
            if len(h["losses"]) > 10:
                h["losses"] = h["losses"][-10:]
                h["sample_counts"] = h["sample_counts"][-10:]

    def _calculate_adaptive_weights(self, updates) -> List[float]:
        base = self._calculate_sample_weights(updates)
        if self.performance_weight == 0:
            return base
        adj = []
        for u in updates:
            h = self.client_performance_history.get(u.client_id)
            if h is None:
                adj.append(1.0)
                continue
            avg_loss = np.mean(h["losses"])
            # upstream takes max() over the per-client loss LISTS (:441-442), i.e. the lexicographically
            # largest history, and then divides by it -- a TypeError for list / float.  The evident intent
            # (largest loss seen) is implemented; with a single-entry history both agree on the value.
            max_loss = max(max(x["losses"]) for x in self.client_performance_history.values() if x["losses"])
            adj.append(1.0 - (avg_loss / max_loss) if max_loss > 0 else 1.0)
        mixed = [(1 - self.performance_weight) * b + self.performance_weight * a for b, a in zip(base, adj)]
        return self._normalize_weights(mixed)


def create_fedavg_aggregator(aggregator_type: str = "standard", **kwargs) -> FedAvgAggregator:
    return AdaptiveFedAvg(**kwargs) if aggregator_type == "adaptive" else FedAvgAggregator(**kwargs)


def benchmark_aggregation_performance(num_clients_list: List[int] = (5, 10, 25, 50), model_size: int = 1000000,
                                      device=None) -> Dict[str, Any]:
    """Same report as the reference's helper (src/aggregation/fedavg.py:487-546: four equal random layers per client, random
    sample counts, `validate_updates=False`) with the client tensors on the CUDA device; `aggregation_time` is the wall
    clock around `aggregate_updates` + a device synchronize.  bench.py / scripts/fedavg_sweep.py are the measured versions."""
    import time
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    results: Dict[str, Any] = {}
    for n in num_clients_list:
        try:
            updates = [ModelUpdate(client_id=f"client_{i}", round_number=1,
                                   model_weights={f"layer{j}": torch.randn(model_size // 4, device=dev) for j in range(1, 5)},
                                   num_samples=int(np.random.randint(100, 1000)), training_loss=float(np.random.uniform(0.1, 2.0)),
                                   privacy_budget_used=0.1, compression_ratio=0.8, timestamp=datetime.now()) for i in range(n)]
            agg = FedAvgAggregator(validate_updates=False)
            torch.cuda.synchronize(dev)
            t0 = time.time()
            model = agg.aggregate_updates(updates)
            torch.cuda.synchronize(dev)
            dt = time.time() - t0
            results[f"{n}_clients"] = {"aggregation_time": dt, "throughput": n / dt,
                                       "memory_usage": sum(w.numel() * w.element_size() for w in model.model_weights.values()),
                                       "participating_clients": len(model.participating_clients)}
        except Exception as e:
            logger.error(f"Benchmark failed for {n} clients: {str(e)}")
            results[f"{n}_clients"] = {"error": str(e)}
    return results
