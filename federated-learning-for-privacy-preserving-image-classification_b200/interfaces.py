"""Abstract contracts of the hot-path classes (the reference's drop-in boundary, src/shared/interfaces.py:75-182).

Only the five contracts the hot path implements are declared -- aggregation, model, data loading, privacy engine,
compression; the coordinator / client *service* interfaces (:18-72) belong to the gRPC orchestration, which is out of
scope.  Method names, argument order and meaning follow the reference so that
``isinstance(obj, AggregationServiceInterface)``-style checks and duck-typed callers (``grpc_server.py:61``,
``round_manager.py:187``, ``federated_trainer.py:127-141``) keep working after the import swap of INTEGRATION.md."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Dict, List

import torch

from .models import CompressedUpdate, GlobalModel, ModelUpdate, ModelWeights

ClientID = str


class AggregationServiceInterface(ABC):
    """interfaces.py:75-96 -- implemented by ``fedavg.FedAvgAggregator``."""

    @abstractmethod
    def aggregate_updates(self, updates: List[ModelUpdate], weights: List[float]) -> GlobalModel: ...

    @abstractmethod
    def validate_update(self, update: ModelUpdate) -> bool: ...

    @abstractmethod
    def compress_global_model(self, model: GlobalModel) -> CompressedUpdate: ...

    @abstractmethod
    def calculate_convergence_metrics(self, old_model: GlobalModel, new_model: GlobalModel) -> float: ...


class ModelInterface(ABC):
    """interfaces.py:99-120 -- implemented by ``models_pytorch.FederatedCNNBase``."""

    @abstractmethod
    def get_model_weights(self) -> ModelWeights: ...

    @abstractmethod
    def set_model_weights(self, weights: ModelWeights) -> None: ...

    @abstractmethod
    def get_parameter_count(self) -> int: ...

    @abstractmethod
    def estimate_memory_usage(self) -> int: ...


class DataLoaderInterface(ABC):
    """interfaces.py:123-139 -- the per-client loader contract (``data_loader.DeviceShardLoader``)."""

    @abstractmethod
    def load_training_data(self, client_id: ClientID) -> torch.utils.data.DataLoader: ...

    @abstractmethod
    def load_validation_data(self) -> torch.utils.data.DataLoader: ...

    @abstractmethod
    def get_data_statistics(self, client_id: ClientID) -> Dict[str, Any]: ...


class PrivacyEngineInterface(ABC):
    """interfaces.py:142-163 -- implemented by ``privacy.DifferentialPrivacyEngine``."""

    @abstractmethod
    def add_noise(self, gradients: ModelWeights, epsilon: float, delta: float) -> ModelWeights: ...

    @abstractmethod
    def clip_gradients(self, gradients: ModelWeights, max_norm: float) -> ModelWeights: ...

    @abstractmethod
    def calculate_privacy_budget(self, epsilon: float, delta: float, steps: int) -> float: ...

    @abstractmethod
    def validate_privacy_parameters(self, epsilon: float, delta: float) -> bool: ...


class CompressionInterface(ABC):
    """interfaces.py:166-182 -- implemented by ``compression.ModelCompressionService``."""

    @abstractmethod
    def compress_weights(self, weights: ModelWeights) -> bytes: ...

    @abstractmethod
    def decompress_weights(self, compressed_data: bytes) -> ModelWeights: ...

    @abstractmethod
    def get_compression_ratio(self, original_size: int, compressed_size: int) -> float: ...
