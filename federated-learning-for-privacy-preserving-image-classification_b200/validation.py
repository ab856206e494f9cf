"""Update validation, same checks and messages as the reference's ``ModelUpdateValidator``
(src/shared/validation.py:21-111) and ``validate_model_compatibility`` (:256-282).  When the update's tensors live on
a CUDA device the three reductions per tensor (isnan / isinf / abs().max(), each a pass and a host sync upstream) are
ONE launch of ``flb_update_stats`` over all layers and one device -> host read."""
from __future__ import annotations

from datetime import datetime, timedelta
from typing import Dict

import torch


class ValidationError(Exception):
    pass


class ModelUpdateValidator:
    def __init__(self, max_weight_magnitude: float = 10.0, min_samples: int = 1):
        self.max_weight_magnitude = max_weight_magnitude
        self.min_samples = min_samples

    def validate_model_update(self, update) -> bool:
        try:
            self._check_fields(update)
            self._check_weights(update.model_weights)
            if not (0 <= update.privacy_budget_used <= 1):
                raise ValidationError("Privacy budget used must be between 0 and 1")
            if not (0 <= update.compression_ratio <= 1):
                raise ValidationError("Compression ratio must be between 0 and 1")
            self._check_timestamp(update.timestamp)
            return True
        except Exception as e:  # validation.py:56-58 wraps everything
            raise ValidationError(f"Model update validation failed: {str(e)}")

    def _check_fields(self, update) -> None:
        if not update.client_id or not isinstance(update.client_id, str):
            raise ValidationError("Client ID must be a non-empty string")
        if update.round_number < 0:
            raise ValidationError("Round number must be non-negative")
        if update.num_samples < self.min_samples:
            raise ValidationError(f"Number of samples must be at least {self.min_samples}")
        if update.training_loss < 0:
            raise ValidationError("Training loss must be non-negative")

    def _check_weights(self, weights: Dict[str, torch.Tensor]) -> None:
        if not weights:
            raise ValidationError("Model weights cannot be empty")
        for name, t in weights.items():
            if not isinstance(t, torch.Tensor):
                raise ValidationError(f"Weight for layer {name} must be a torch.Tensor")
        ts = list(weights.values())
        if all(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.device == ts[0].device for t in ts):
            from . import ops
            dev = ts[0].device
            offs = [0]
            for t in ts:
                offs.append(offs[-1] + t.numel())
            table = torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64).to(dev)
            mx, fl = ops.update_stats(table, torch.tensor(offs, dtype=torch.int64, device=dev), 1, offs[-1], dev)
            mx, fl = mx[0].cpu().tolist(), fl[0].cpu().tolist()          # the one read
            for name, m, f in zip(weights, mx, fl):
                if f & 1:
                    raise ValidationError(f"NaN values found in layer {name}")
                if f & 2:
                    raise ValidationError(f"Infinite values found in layer {name}")
                if m > self.max_weight_magnitude:
                    raise ValidationError(
                        f"Weight magnitude {m} exceeds maximum {self.max_weight_magnitude} in layer {name}")
            return
        for name, t in weights.items():
            if torch.isnan(t).any():
                raise ValidationError(f"NaN values found in layer {name}")
            if torch.isinf(t).any():
                raise ValidationError(f"Infinite values found in layer {name}")
            mag = torch.abs(t).max().item()
            if mag > self.max_weight_magnitude:
                raise ValidationError(
                    f"Weight magnitude {mag} exceeds maximum {self.max_weight_magnitude} in layer {name}")

    def _check_timestamp(self, ts: datetime) -> None:
        now = datetime.now()
        if ts < now - timedelta(hours=24):
            raise ValidationError("Model update timestamp is too old")
        if ts > now + timedelta(minutes=5):
            raise ValidationError("Model update timestamp is in the future")


def validate_model_compatibility(model1_weights: Dict[str, torch.Tensor],
                                 model2_weights: Dict[str, torch.Tensor]) -> bool:
    if set(model1_weights.keys()) != set(model2_weights.keys()):
        raise ValidationError("Models have different layer names")
    for name in model1_weights:
        if model1_weights[name].shape != model2_weights[name].shape:
            raise ValidationError(f"Layer {name} has incompatible shapes: "
                                  f"{model1_weights[name].shape} vs {model2_weights[name].shape}")
    return True
