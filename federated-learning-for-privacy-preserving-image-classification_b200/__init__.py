"""flb200 -- B200-native (sm_100a) federated client-training / DP / FedAvg hot path.

Drop-in surface (same call signatures as the reference):
    ModelFactory, SimpleCNN, CIFAR10CNN          (reference src/shared/models_pytorch.py)
    LocalTrainer                                  (reference src/shared/training.py)
    DifferentialPrivacyEngine, create_privacy_engine   (reference src/shared/privacy.py)
    FedAvgAggregator, AdaptiveFedAvg              (reference src/aggregation/fedavg.py)
    ModelCompressionService                       (reference src/shared/compression.py)
    SimulationConfig, FederatedLearningSimulation (reference src/simulation/federated_simulation.py)
All arithmetic runs in hand-written CUDA kernels behind the C ABI in ``include/flb.h`` (``libflb.so``);
there is no CPU fallback -- a missing library or GPU raises."""
from ._lib import FlbError, LIB_PATH, load as load_library  # noqa: F401
from .models import (CompressedUpdate, GlobalModel, ModelUpdate, ModelWeights, PrivacyConfig, RoundConfig,  # noqa: F401
                     TrainingMetrics)

__all__ = ["FlbError", "load_library", "ModelUpdate", "GlobalModel", "PrivacyConfig", "TrainingMetrics"]
