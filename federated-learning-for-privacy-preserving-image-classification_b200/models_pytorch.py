"""Model definitions on the hot path, same names / parameter names / shapes / init as the reference's
``src/shared/models_pytorch.py`` (FederatedCNNBase :18-56, SimpleCNN :59-97, CIFAR10CNN :100-165,
ModelFactory :331-424).  The modules are parameter containers built from stock ``nn`` layers so that
``state_dict`` keys, ``named_parameters()`` order and default initialisation match the reference exactly;
training never calls their ``forward`` -- ``LocalTrainer`` / ``BatchedClientTrainer`` run the CUDA kernels on
the flattened parameters.  ``forward`` (inference) is routed through the same kernels.
FederatedResNet / LightweightMobileNet are outside the hot-path scope."""
from __future__ import annotations

import logging
from collections import OrderedDict
from typing import Any, Dict, List

import torch
import torch.nn as nn

from .interfaces import ModelInterface
from .models import ModelWeights

logger = logging.getLogger(__name__)

MODEL_IDS = {"simple_cnn": 0, "cifar10_cnn": 1}
INPUT_SHAPES = {"simple_cnn": (1, 28, 28), "cifar10_cnn": (3, 32, 32)}


class FederatedCNNBase(nn.Module, ModelInterface):
    def __init__(self):
        super().__init__()
        self.model_name = "base_cnn"

    def get_model_weights(self) -> ModelWeights:
        """Fresh clones of the parameters only -- buffers are not federated (models_pytorch.py:25-27)."""
        return {name: p.data.clone() for name, p in self.named_parameters()}

    def set_model_weights(self, weights: ModelWeights) -> None:
        sd = self.state_dict()
        for name, w in weights.items():
            if name in sd:
                sd[name].copy_(w)
            else:
                logger.warning(f"Weight {name} not found in model state dict")

    def get_parameter_count(self) -> int:
        return sum(p.numel() for p in self.parameters())

    def estimate_memory_usage(self) -> int:
        return (sum(p.numel() * p.element_size() for p in self.parameters())
                + sum(b.numel() * b.element_size() for b in self.buffers()))

    def get_model_info(self) -> Dict[str, Any]:
        return {"name": self.model_name, "parameters": self.get_parameter_count(),
                "memory_bytes": self.estimate_memory_usage(), "layers": len(list(self.named_modules())),
                "trainable_params": sum(p.numel() for p in self.parameters() if p.requires_grad)}

    def param_spec(self) -> "OrderedDict[str, tuple]":
        return OrderedDict((n, tuple(p.shape)) for n, p in self.named_parameters())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from .training import forward_logits           # late import: training imports this module
        return forward_logits(self, x)


class SimpleCNN(FederatedCNNBase):
    """28x28x1 -> conv3x3(32)+ReLU+pool -> conv3x3(64)+ReLU+pool -> fc128+ReLU+dropout -> fc(num_classes)."""

    def __init__(self, num_classes: int = 10, dropout_rate: float = 0.25):
        super().__init__()
        self.model_name = "simple_cnn"
        self.num_classes = num_classes
        self.dropout_rate = dropout_rate
        self.conv1 = nn.Conv2d(1, 32, kernel_size=3, stride=1, padding=1)
        self.conv2 = nn.Conv2d(32, 64, kernel_size=3, stride=1, padding=1)
        self.pool = nn.MaxPool2d(kernel_size=2, stride=2)
        self.dropout = nn.Dropout(dropout_rate)
        self.fc1 = nn.Linear(64 * 7 * 7, 128)
        self.fc2 = nn.Linear(128, num_classes)


class CIFAR10CNN(FederatedCNNBase):
    """32x32x3 -> 3 x [conv-bn-relu, conv-bn-relu, pool, dropout] -> fc512 -> fc256 -> fc(num_classes)."""

    def __init__(self, num_classes: int = 10, dropout_rate: float = 0.3):
        super().__init__()
        self.model_name = "cifar10_cnn"
        self.num_classes = num_classes
        self.dropout_rate = dropout_rate
        chans = [(3, 32), (32, 32), (32, 64), (64, 64), (64, 128), (128, 128)]
        for i, (cin, cout) in enumerate(chans, start=1):
            setattr(self, f"conv{i}", nn.Conv2d(cin, cout, kernel_size=3, padding=1))
            setattr(self, f"bn{i}", nn.BatchNorm2d(cout))
        self.pool = nn.MaxPool2d(2, 2)
        self.dropout = nn.Dropout(dropout_rate)
        self.fc1 = nn.Linear(128 * 4 * 4, 512)
        self.fc2 = nn.Linear(512, 256)
        self.fc3 = nn.Linear(256, num_classes)


class ModelFactory:
    AVAILABLE_MODELS = {"simple_cnn": SimpleCNN, "cifar10_cnn": CIFAR10CNN}

    @classmethod
    def create_model(cls, model_name: str, **kwargs) -> FederatedCNNBase:
        if model_name not in cls.AVAILABLE_MODELS:
            raise ValueError(f"Unknown model: {model_name}. Available: {list(cls.AVAILABLE_MODELS.keys())}")
        return cls.AVAILABLE_MODELS[model_name](**kwargs)

    @classmethod
    def get_model_for_dataset(cls, dataset: str, **kwargs) -> FederatedCNNBase:
        dataset = dataset.lower()
        if dataset == "mnist":
            return cls.create_model("simple_cnn", num_classes=10, **kwargs)
        if dataset == "cifar10":
            return cls.create_model("cifar10_cnn", num_classes=10, **kwargs)
        logger.warning(f"Unknown dataset {dataset}, using simple CNN")
        return cls.create_model("simple_cnn", **kwargs)

    @classmethod
    def list_available_models(cls) -> List[str]:
        return list(cls.AVAILABLE_MODELS.keys())

    @classmethod
    def get_model_info(cls, model_name: str) -> Dict[str, Any]:
        if model_name not in cls.AVAILABLE_MODELS:
            raise ValueError(f"Unknown model: {model_name}")
        return cls.create_model(model_name).get_model_info()
