"""Round-to-round convergence tracking with the reference's ``ConvergenceDetector`` surface
(``src/aggregation/convergence.py``: ConvergenceMetrics :24-34, ConvergenceDetector :37-335, AdaptiveConvergenceDetector
:338-398, create_convergence_detector :401-415) -- SURVEY.md section 8(f) row 1.

The arithmetic on the path is ``_calculate_weight_change_metrics`` (:189-217): ||theta_new - theta_old||_2 and ||theta_new||_2
over all layers.  Upstream that is two ``torch.norm(...).item()`` calls (two passes and two host syncs) per layer; here, for
CUDA tensors, it is ONE launch of ``flb_delta_norms`` over all layers (per-layer sums of squares in double) and one read.
Everything else is host bookkeeping over a few floats per round and keeps the reference's fields, thresholds and messages.
Not mirrored: the offline ``analyze_convergence_patterns`` report helper (:418-466)."""
from __future__ import annotations

import logging
import math
from collections import deque
from datetime import datetime
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from .models import GlobalModel, ModelWeights

logger = logging.getLogger(__name__)


class ConvergenceError(Exception):
    pass


class ConvergenceMetrics:
    def __init__(self):
        self.weight_change_norm = self.relative_weight_change = 0.0
        self.accuracy_change = self.loss_change = 0.0
        self.convergence_score = 0.0
        self.is_converged = False
        self.confidence = 0.0


def _sumsq_pairs(cur: ModelWeights, prev: ModelWeights) -> Tuple[float, float]:
    """(sum ||cur_l - prev_l||^2, sum ||cur_l||^2) over the layers present in both."""
    names = [n for n in cur if n in prev]
    if not names:
        return 0.0, 0.0
    c, p = [cur[n] for n in names], [prev[n] for n in names]
    fused = all(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.device == c[0].device and t.shape == o.shape
                for t, o in zip(c + p, p + c))
    if fused:
        from . import ops
        dev, offs = c[0].device, [0]
        for t in c:
            offs.append(offs[-1] + t.numel())
        s = ops.delta_norms(torch.tensor([t.data_ptr() for t in c], dtype=torch.int64).to(dev),
                            torch.tensor([t.data_ptr() for t in p], dtype=torch.int64).to(dev),
                            torch.tensor(offs, dtype=torch.int64, device=dev), offs[-1], dev).sum(0).cpu().tolist()
        return float(s[0]), float(s[1])
    d2 = n2 = 0.0
    for a, b in zip(c, p):                                   # host tensors: the reference's own formula
        d2 += torch.norm(a - b.to(a.device)).item() ** 2
        n2 += torch.norm(a).item() ** 2
    return d2, n2


class ConvergenceDetector:
    def __init__(self, patience: int = 5, min_delta: float = 1e-4, window_size: int = 3, convergence_threshold: float = 1e-3):
        self.patience, self.min_delta = patience, min_delta
        self.window_size, self.convergence_threshold = window_size, convergence_threshold
        self.accuracy_history, self.loss_history = deque(maxlen=100), deque(maxlen=100)
        self.weight_change_history, self.convergence_history = deque(maxlen=100), deque(maxlen=100)
        self.best_accuracy, self.best_loss = 0.0, float("inf")
        self.rounds_without_improvement = 0
        self.converged = False

    # ---- the per-round entry (convergence.py:73-150) ---------------------------------------------------------------
    def calculate_convergence_metrics(self, current_model: GlobalModel,
                                      previous_model: Optional[GlobalModel] = None) -> ConvergenceMetrics:
        try:
            m = ConvergenceMetrics()
            acc = current_model.get_accuracy() or 0.0
            loss = self._extract_loss_from_model(current_model)
            self.accuracy_history.append(acc)
            self.loss_history.append(loss)
            if previous_model is not None:
                wc = self._calculate_weight_change_metrics(current_model.model_weights, previous_model.model_weights)
                m.weight_change_norm, m.relative_weight_change = wc["norm"], wc["relative"]
                m.accuracy_change = acc - (previous_model.get_accuracy() or 0.0)
                m.loss_change = loss - self._extract_loss_from_model(previous_model)
                self.weight_change_history.append(m.weight_change_norm)
            m.convergence_score = self._calculate_convergence_score(m)
            m.is_converged, m.confidence = self._check_convergence(m)
            self.convergence_history.append({"round": current_model.round_number, "accuracy": acc, "loss": loss,
                                             "convergence_score": m.convergence_score, "is_converged": m.is_converged,
                                             "timestamp": datetime.now().isoformat()})
            if acc > self.best_accuracy:
                self.best_accuracy, self.rounds_without_improvement = acc, 0
            else:
                self.rounds_without_improvement += 1
            self.best_loss = min(self.best_loss, loss)
            return m
        except Exception as e:
            logger.error(f"Convergence calculation failed: {str(e)}")
            raise ConvergenceError(f"Convergence calculation failed: {str(e)}")

    def should_stop_early(self) -> Tuple[bool, str]:
        try:
            if self.rounds_without_improvement >= self.patience:
                return True, f"No improvement for {self.patience} rounds"
            w = self.window_size
            if len(self.convergence_history) >= w:
                avg = float(np.mean([h["convergence_score"] for h in list(self.convergence_history)[-w:]]))
                if avg < self.convergence_threshold:
                    return True, f"Convergence threshold reached (score: {avg:.6f})"
            if len(self.accuracy_history) >= 2 * w:
                hist = list(self.accuracy_history)
                change = abs(float(np.mean(hist[-w:])) - float(np.mean(hist[-2 * w:-w])))
                if change < self.min_delta:
                    return True, f"Accuracy plateaued (change: {change:.6f})"
            return False, "Continue training"
        except Exception as e:
            logger.error(f"Early stopping check failed: {str(e)}")
            return False, "Error in early stopping check"

    # ---- the arithmetic (convergence.py:189-217) -------------------------------------------------------------------
    def _calculate_weight_change_metrics(self, current_weights: ModelWeights, previous_weights: ModelWeights) -> Dict[str, float]:
        d2, n2 = _sumsq_pairs(current_weights, previous_weights)
        change, total = math.sqrt(d2), math.sqrt(n2)
        return {"norm": change, "relative": change / total if total > 0 else 0.0}

    def _extract_loss_from_model(self, model: GlobalModel) -> float:
        for key, value in model.accuracy_metrics.items():
            if "loss" in key.lower():
                return float(value)
        return 0.0

    def _calculate_convergence_score(self, m: ConvergenceMetrics) -> float:      # lower = closer to converged
        return max(m.relative_weight_change, 0.0) + max(-m.accuracy_change, 0.0) + max(m.loss_change, 0.0)

    def _check_convergence(self, m: ConvergenceMetrics) -> Tuple[bool, float]:
        confidence = 0.0
        if len(self.convergence_history) >= 3:
            last = [h["convergence_score"] for h in list(self.convergence_history)[-3:]]
            if np.mean(last) < self.convergence_threshold:
                confidence = max(0.0, 1.0 - float(np.std(last)))
        return m.convergence_score < self.convergence_threshold, confidence

    def get_convergence_summary(self) -> Dict[str, Any]:
        if not self.convergence_history:
            return {"message": "No convergence data available"}
        recent = list(self.convergence_history)[-10:]
        stop, reason = self.should_stop_early()
        return {"current_status": {"converged": self.converged, "best_accuracy": self.best_accuracy, "best_loss": self.best_loss,
                                   "rounds_without_improvement": self.rounds_without_improvement,
                                   "total_rounds": len(self.convergence_history)},
                "recent_performance": {"avg_accuracy": np.mean([h["accuracy"] for h in recent]),
                                       "avg_loss": np.mean([h["loss"] for h in recent]),
                                       "avg_convergence_score": np.mean([h["convergence_score"] for h in recent]),
                                       "convergence_trend": self._calculate_trend([h["convergence_score"] for h in recent])},
                "early_stopping": {"patience": self.patience, "min_delta": self.min_delta, "should_stop": stop, "stop_reason": reason}}

    def _calculate_trend(self, values: List[float]) -> str:
        if len(values) < 2:
            return "insufficient_data"
        slope = np.polyfit(np.arange(len(values)), values, 1)[0]
        return "improving" if slope < -0.001 else ("degrading" if slope > 0.001 else "stable")

    def reset(self):
        for h in (self.accuracy_history, self.loss_history, self.weight_change_history, self.convergence_history):
            h.clear()
        self.best_accuracy, self.best_loss = 0.0, float("inf")
        self.rounds_without_improvement = 0
        self.converged = False


class AdaptiveConvergenceDetector(ConvergenceDetector):
    """Threshold drifts with the variance of the last five scores, inside [0.5, 2] x the initial value (:338-398)."""

    def __init__(self, patience: int = 5, min_delta: float = 1e-4, window_size: int = 3, convergence_threshold: float = 1e-3,
                 adaptation_rate: float = 0.1):
        super().__init__(patience, min_delta, window_size, convergence_threshold)
        self.initial_threshold, self.adaptation_rate = convergence_threshold, adaptation_rate

    def calculate_convergence_metrics(self, current_model, previous_model=None) -> ConvergenceMetrics:
        m = super().calculate_convergence_metrics(current_model, previous_model)
        self._adapt_threshold()
        return m

    def _adapt_threshold(self):
        if len(self.convergence_history) < 5:
            return
        var = float(np.var([h["convergence_score"] for h in list(self.convergence_history)[-5:]]))
        if var > 0.01:
            self.convergence_threshold = min(self.initial_threshold * 2, self.convergence_threshold * (1 + self.adaptation_rate))
        elif var < 0.001:
            self.convergence_threshold = max(self.initial_threshold * 0.5, self.convergence_threshold * (1 - self.adaptation_rate))


def create_convergence_detector(detector_type: str = "standard", **kwargs) -> ConvergenceDetector:
    return AdaptiveConvergenceDetector(**kwargs) if detector_type == "adaptive" else ConvergenceDetector(**kwargs)
