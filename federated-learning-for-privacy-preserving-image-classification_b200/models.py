"""Value types that cross the hot-path boundary, field-for-field with the reference's
``src/shared/models.py`` (PrivacyConfig :20-37, ModelUpdate :50-72, GlobalModel :75-87,
TrainingMetrics :90-97, RoundConfig :126-135, CompressedUpdate :149-164, aliases :168-170) so that
objects built by reference orchestration code can be handed to the classes in this package unchanged
(only attribute access is used -- the reference's own dataclass instances work as well)."""
from __future__ import annotations

from dataclasses import dataclass
from datetime import datetime
from typing import Any, Dict, List, Optional

import torch

ModelWeights = Dict[str, torch.Tensor]
ClientID = str
RoundNumber = int


@dataclass
class PrivacyConfig:
    epsilon: float
    delta: float
    max_grad_norm: float
    noise_multiplier: float

    def __post_init__(self):
        # same four checks and messages as models.py:27-37
        if self.epsilon <= 0:
            raise ValueError("Epsilon must be positive")
        if self.delta < 0 or self.delta >= 1:
            raise ValueError("Delta must be in [0, 1)")
        if self.max_grad_norm <= 0:
            raise ValueError("Max gradient norm must be positive")
        if self.noise_multiplier < 0:
            raise ValueError("Noise multiplier must be non-negative")


@dataclass
class ModelUpdate:
    client_id: str
    round_number: int
    model_weights: ModelWeights
    num_samples: int
    training_loss: float
    privacy_budget_used: float
    compression_ratio: float
    timestamp: datetime

    def validate(self) -> bool:
        ok = bool(self.client_id) and self.round_number >= 0
        ok = ok and self.num_samples > 0 and self.training_loss >= 0
        ok = ok and 0 <= self.privacy_budget_used <= 1 and 0 <= self.compression_ratio <= 1
        return ok


@dataclass
class GlobalModel:
    round_number: int
    model_weights: ModelWeights
    accuracy_metrics: Dict[str, float]
    participating_clients: List[str]
    convergence_score: float
    created_at: datetime

    def get_accuracy(self, dataset: str = "test") -> Optional[float]:
        return self.accuracy_metrics.get(f"{dataset}_accuracy")


@dataclass
class TrainingMetrics:
    loss: float
    accuracy: float
    epochs_completed: int
    training_time: float
    samples_processed: int


@dataclass
class RoundConfig:
    round_number: int
    min_clients: int
    max_clients: int
    local_epochs: int
    batch_size: int
    learning_rate: float
    timeout_seconds: int


@dataclass
class CompressedUpdate:
    client_id: str
    round_number: int
    compressed_weights: bytes
    compression_metadata: Dict[str, Any]
    original_size: int
    compressed_size: int

    @property
    def compression_ratio(self) -> float:
        return 0.0 if self.original_size == 0 else self.compressed_size / self.original_size
