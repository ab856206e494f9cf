"""Update-level differential privacy behind the reference's call signatures.

Mirrors ``src/shared/privacy.py``: ``PrivacyBudgetTracker`` (:25-92), ``GradientClipper`` (:95-168),
``GaussianNoiseGenerator`` (:171-254), ``DifferentialPrivacyEngine`` (:257-416), ``PrivacyAccountant``
(:419-484), ``create_privacy_engine`` (:487-512).  The bookkeeping (budget, ledger, parameter checks,
exception types and messages) is host Python with the reference's semantics; the arithmetic -- global L2
norm over all tensors, clip, Gaussian noise -- runs in ``csrc/privacy.cu`` on a flat fp32 row.  When the
stock clipper and noise generator are attached, ``add_noise`` is one fused pair of launches; replacing
``engine.noise_generator`` (the reference's test-injection point) falls back to calling
the two objects in sequence exactly like privacy.py:295-301."""
from __future__ import annotations

import json
import logging
import math
from collections import OrderedDict
from datetime import datetime
from typing import Any, Dict, List, Optional, Tuple

import torch

from . import _lib as L
from . import ops
from .interfaces import PrivacyEngineInterface
from .layout import ParamLayout
from .models import ModelWeights, PrivacyConfig

logger = logging.getLogger(__name__)


class PrivacyError(Exception):
    pass


def _default_device(device) -> torch.device:
    if device is None:
        return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cuda")
    return torch.device(device)


def _pack(gradients: ModelWeights, device: torch.device):
    """dict (entries may be None, any device) -> (layout over the non-None entries, [1, ld] device row)."""
    present = OrderedDict((k, v) for k, v in gradients.items() if v is not None)
    layout = ParamLayout.from_weights(present)
    row = layout.new_rows(1, device)
    layout.flatten_into(row[0], {k: v.detach().to(torch.float32) for k, v in present.items()})
    return layout, row


def _unpack(layout: ParamLayout, row: torch.Tensor, like: ModelWeights) -> ModelWeights:
    out: ModelWeights = {}
    views = layout.views(row)
    for k, v in like.items():
        out[k] = None if v is None else views[k].to(v.device, copy=True)
    return out


class PrivacyBudgetTracker:
    """Additive (epsilon, delta) ledger; privacy.py:25-92."""

    def __init__(self, initial_epsilon: float, initial_delta: float):
        self.initial_epsilon = initial_epsilon
        self.initial_delta = initial_delta
        self.consumed_epsilon = 0.0
        self.consumed_delta = 0.0
        self.consumption_history: List[Dict[str, Any]] = []
        self.start_time = datetime.now()

    def consume_budget(self, epsilon: float, delta: float, operation: str = "training"):
        self.consumed_epsilon += epsilon
        self.consumed_delta += delta
        self.consumption_history.append({
            "timestamp": datetime.now().isoformat(), "epsilon": epsilon, "delta": delta, "operation": operation,
            "total_epsilon": self.consumed_epsilon, "total_delta": self.consumed_delta})

    def get_remaining_budget(self) -> Tuple[float, float]:
        return (max(0, self.initial_epsilon - self.consumed_epsilon),
                max(0, self.initial_delta - self.consumed_delta))

    def is_budget_exhausted(self, required_epsilon: float = 0, required_delta: float = 0) -> bool:
        eps, dlt = self.get_remaining_budget()
        return eps < required_epsilon or dlt < required_delta

    def get_budget_status(self) -> Dict[str, Any]:
        eps, dlt = self.get_remaining_budget()
        return {
            "initial_epsilon": self.initial_epsilon, "initial_delta": self.initial_delta,
            "consumed_epsilon": self.consumed_epsilon, "consumed_delta": self.consumed_delta,
            "remaining_epsilon": eps, "remaining_delta": dlt,
            "epsilon_utilization": self.consumed_epsilon / self.initial_epsilon,
            "delta_utilization": self.consumed_delta / self.initial_delta,
            "operations_count": len(self.consumption_history),
            "tracking_duration": (datetime.now() - self.start_time).total_seconds()}


class GradientClipper:
    """Global-L2 clip over all tensors of a dict; privacy.py:107-144."""

    def __init__(self, max_grad_norm: float, device: Optional[torch.device] = None):
        self.max_grad_norm = max_grad_norm
        self.device = _default_device(device)

    def clip_gradients(self, gradients: ModelWeights) -> Tuple[ModelWeights, float]:
        try:
            layout, row = _pack(gradients, self.device)
            out, norms = ops.dp_clip_noise(row, None, self.max_grad_norm, 0.0, P=layout.P)
            total = float(norms.item())                      # the one host sync (reference: one per tensor)
            return _unpack(layout, out[0], gradients), min(total, self.max_grad_norm)
        except Exception as e:
            raise PrivacyError(f"Gradient clipping failed: {str(e)}")

    def estimate_sensitivity(self, gradients_batch: List[ModelWeights]) -> float:
        best = 0.0
        for g in gradients_batch or []:
            layout, row = _pack(g, self.device)
            _, norms = ops.dp_clip_noise(row, None, 1.0, 0.0, P=layout.P)
            best = max(best, float(norms.item()))
        return best


class GaussianNoiseGenerator:
    """sigma = S*sqrt(2 ln(1.25/delta))/eps (privacy.py:209); normals from Philox4x32-10 + Box-Muller in
    registers instead of ``torch.normal`` (:212).  Every call advances ``self.stream`` so draws never repeat, and the
    seed defaults to fresh OS entropy per generator (like torch's global RNG upstream): two engines, or two client
    processes, never share a noise sequence.  ``seed=`` is for reproducible tests only."""

    def __init__(self, device: Optional[torch.device] = None, seed: Optional[int] = None):
        self.device = _default_device(device)
        self.seed = L.fresh_seed() if seed is None else int(seed) & (2**64 - 1)
        self.stream = 0

    @staticmethod
    def sigma(sensitivity: float, epsilon: float, delta: float) -> float:
        return sensitivity * ops.gaussian_sigma_unit(epsilon, delta)

    def generate_noise(self, shape, sensitivity: float, epsilon: float, delta: float) -> torch.Tensor:
        try:
            sigma = self.sigma(sensitivity, epsilon, delta)
            n = 1
            for s in shape:
                n *= int(s)
            z = ops.philox_normal(n, self.seed, self.stream, self.device)
            self.stream += 1
            return (z * sigma).view(tuple(shape))
        except Exception as e:
            raise PrivacyError(f"Noise generation failed: {str(e)}")

    def add_noise_to_gradients(self, gradients: ModelWeights, sensitivity: float, epsilon: float,
                               delta: float) -> ModelWeights:
        try:
            sigma = self.sigma(sensitivity, epsilon, delta)
            layout, row = _pack(gradients, self.device)
            out = ops.dp_add_noise(row, sigma, self.seed, self.stream, P=layout.P)
            self.stream += 1
            return _unpack(layout, out[0], gradients)
        except Exception as e:
            raise PrivacyError(f"Adding noise to gradients failed: {str(e)}")


class DifferentialPrivacyEngine(PrivacyEngineInterface):
    """privacy.py:257-416.  ``device=None`` means the current CUDA device (the reference defaults to CPU;
    this package has no CPU path).  Inputs may live on the host: they are staged to the device and the
    result is returned on the input's device, as new tensors (inputs are never mutated)."""

    def __init__(self, privacy_config: PrivacyConfig, device: Optional[torch.device] = None, seed: Optional[int] = None):
        self.config = privacy_config
        self.device = _default_device(device)
        self.clipper = GradientClipper(privacy_config.max_grad_norm, self.device)
        self.noise_generator = GaussianNoiseGenerator(self.device, seed)      # seed=None: fresh entropy
        self.budget_tracker = PrivacyBudgetTracker(privacy_config.epsilon, privacy_config.delta)

    # -- the hot call ----------------------------------------------------------------------------
    def add_noise(self, gradients: ModelWeights, epsilon: float, delta: float) -> ModelWeights:
        try:
            if not self.validate_privacy_parameters(epsilon, delta):
                raise PrivacyError("Invalid privacy parameters")
            if self.budget_tracker.is_budget_exhausted(epsilon, delta):
                raise PrivacyError("Privacy budget exhausted")
            if type(self.clipper) is GradientClipper and type(self.noise_generator) is GaussianNoiseGenerator:
                noisy = self._fused_clip_noise(gradients, epsilon, delta)
            else:
                clipped, actual_norm = self.clipper.clip_gradients(gradients)
                noisy = self.noise_generator.add_noise_to_gradients(clipped, actual_norm, epsilon, delta)
            self.budget_tracker.consume_budget(epsilon, delta, "gradient_noise")
            return noisy
        except Exception as e:
            logger.error(f"Adding DP noise failed: {str(e)}")
            raise PrivacyError(f"Adding DP noise failed: {str(e)}")

    def _fused_clip_noise(self, gradients: ModelWeights, epsilon: float, delta: float) -> ModelWeights:
        layout, row = _pack(gradients, self.device)
        gen = self.noise_generator
        out, _ = ops.dp_clip_noise(row, None, self.clipper.max_grad_norm, ops.gaussian_sigma_unit(epsilon, delta),
                                   seed=gen.seed, stream_base=gen.stream, P=layout.P)
        gen.stream += 1
        return _unpack(layout, out[0], gradients)

    def add_noise_rows(self, local: torch.Tensor, global_row: Optional[torch.Tensor], epsilon: float, delta: float,
                       P: int, z: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
        """Batched client glue (src/client/federated_trainer.py:428-469) for K clients at once:
        rows_out[k] = global + noise(clip(local[k] - global)).  One budget charge, K Philox streams.
        Returns (rows_out [K, ld], ||delta_k|| [K] device tensor)."""
        if not self.validate_privacy_parameters(epsilon, delta):
            raise PrivacyError("Adding DP noise failed: Invalid privacy parameters")
        if self.budget_tracker.is_budget_exhausted(epsilon, delta):
            raise PrivacyError("Adding DP noise failed: Privacy budget exhausted")
        gen = self.noise_generator
        res = ops.dp_clip_noise(local, global_row, self.config.max_grad_norm, ops.gaussian_sigma_unit(epsilon, delta),
                                seed=gen.seed, stream_base=gen.stream, z=z, P=P, out=out)
        gen.stream += local.shape[0]
        self.budget_tracker.consume_budget(epsilon, delta, "gradient_noise")
        return res

    def clip_gradients(self, gradients: ModelWeights, max_norm: float) -> ModelWeights:
        clipped, _ = GradientClipper(max_norm, self.device).clip_gradients(gradients)
        return clipped

    # -- bookkeeping (host) ------------------------------------------------------------------------
    def calculate_privacy_budget(self, epsilon: float, delta: float, steps: int) -> float:
        if steps <= 1:
            return epsilon
        return epsilon * math.sqrt(2 * steps * math.log(1 / delta)) + steps * epsilon * (math.exp(epsilon) - 1)

    def validate_privacy_parameters(self, epsilon: float, delta: float) -> bool:
        try:
            if epsilon <= 0:
                logger.error("Epsilon must be positive")
                return False
            if epsilon > 10.0:
                logger.warning(f"Epsilon {epsilon} is very high, privacy may be weak")
            if delta <= 0 or delta >= 1:
                logger.error("Delta must be in (0, 1)")
                return False
            if delta > 1e-3:
                logger.warning(f"Delta {delta} is high, privacy may be weak")
            return True
        except Exception:
            return False

    def get_privacy_analysis(self) -> Dict[str, Any]:
        order = ["strong", "moderate", "weak"]
        e = "strong" if self.config.epsilon < 1.0 else "moderate" if self.config.epsilon < 5.0 else "weak"
        d = "strong" if self.config.delta < 1e-5 else "moderate" if self.config.delta < 1e-3 else "weak"
        return {
            "privacy_config": {"epsilon": self.config.epsilon, "delta": self.config.delta,
                               "max_grad_norm": self.config.max_grad_norm,
                               "noise_multiplier": self.config.noise_multiplier},
            "budget_status": self.budget_tracker.get_budget_status(),
            "privacy_strength": {"epsilon_strength": e, "delta_strength": d,
                                 "overall_strength": order[max(order.index(e), order.index(d))]},
            "recommendations": self._recommendations()}

    def _recommendations(self) -> List[str]:
        rec = []
        if self.config.epsilon > 5.0:
            rec.append("Consider reducing epsilon for stronger privacy")
        if self.config.delta > 1e-3:
            rec.append("Consider reducing delta for better privacy guarantees")
        if self.config.max_grad_norm > 10.0:
            rec.append("Consider reducing gradient clipping norm to improve privacy")
        if self.budget_tracker.get_remaining_budget()[0] < self.config.epsilon * 0.1:
            rec.append("Privacy budget nearly exhausted, consider resetting or reducing usage")
        return rec or ["Privacy configuration looks good"]

    def reset_budget(self, new_epsilon: Optional[float] = None, new_delta: Optional[float] = None):
        self.budget_tracker = PrivacyBudgetTracker(new_epsilon or self.config.epsilon, new_delta or self.config.delta)
        if new_epsilon:
            self.config.epsilon = new_epsilon
        if new_delta:
            self.config.delta = new_delta


class PrivacyAccountant:
    """Plain additive ledger; privacy.py:419-484."""

    def __init__(self):
        self.privacy_ledger: List[Dict[str, Any]] = []
        self.total_epsilon = 0.0
        self.total_delta = 0.0

    def add_mechanism(self, mechanism_type: str, epsilon: float, delta: float, sensitivity: float,
                      noise_scale: float, metadata: Optional[Dict[str, Any]] = None):
        self.privacy_ledger.append({"timestamp": datetime.now().isoformat(), "mechanism_type": mechanism_type,
                                    "epsilon": epsilon, "delta": delta, "sensitivity": sensitivity,
                                    "noise_scale": noise_scale, "metadata": metadata or {}})
        self.total_epsilon += epsilon
        self.total_delta += delta

    def get_total_privacy_cost(self) -> Tuple[float, float]:
        return self.total_epsilon, self.total_delta

    def get_privacy_ledger(self) -> List[Dict[str, Any]]:
        return self.privacy_ledger.copy()

    def export_ledger(self, filepath: str):
        try:
            with open(filepath, "w") as f:
                json.dump({"total_epsilon": self.total_epsilon, "total_delta": self.total_delta,
                           "ledger": self.privacy_ledger}, f, indent=2)
        except Exception as e:
            logger.error(f"Failed to export privacy ledger: {str(e)}")


def create_privacy_engine(epsilon: float = 1.0, delta: float = 1e-5, max_grad_norm: float = 1.0,
                          noise_multiplier: float = 1.0,
                          device: Optional[torch.device] = None, seed: Optional[int] = None) -> DifferentialPrivacyEngine:
    return DifferentialPrivacyEngine(PrivacyConfig(epsilon=epsilon, delta=delta, max_grad_norm=max_grad_norm,
                                                   noise_multiplier=noise_multiplier), device, seed)


def estimate_privacy_parameters(target_accuracy: float = 0.9, dataset_size: int = 10000,
                                num_rounds: int = 100) -> Dict[str, float]:
    """Host heuristic, privacy.py:515-556."""
    eps = 1.0 if dataset_size > 5000 else 2.0
    if target_accuracy > 0.95:
        eps *= 2
    elif target_accuracy < 0.85:
        eps *= 0.5
    return {"epsilon": eps / math.sqrt(num_rounds), "delta": 1.0 / dataset_size,
            "max_grad_norm": 1.0 if target_accuracy > 0.9 else 2.0, "noise_multiplier": 1.0}
