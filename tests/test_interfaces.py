"""CPU: the reference's abstract contracts (src/shared/interfaces.py:75-182) are declared and implemented."""
import inspect

import flb200  # noqa: F401
from flb200 import compression, data_loader, fedavg, interfaces, models_pytorch, privacy


def test_hot_path_classes_implement_the_reference_contracts():
    assert issubclass(fedavg.FedAvgAggregator, interfaces.AggregationServiceInterface)
    assert issubclass(fedavg.AdaptiveFedAvg, interfaces.AggregationServiceInterface)
    assert issubclass(models_pytorch.FederatedCNNBase, interfaces.ModelInterface)
    assert issubclass(privacy.DifferentialPrivacyEngine, interfaces.PrivacyEngineInterface)
    assert issubclass(compression.ModelCompressionService, interfaces.CompressionInterface)
    assert issubclass(data_loader.DeviceShardLoader, interfaces.DataLoaderInterface)
    for cls in (fedavg.FedAvgAggregator, models_pytorch.SimpleCNN, models_pytorch.CIFAR10CNN, privacy.DifferentialPrivacyEngine,
                compression.ModelCompressionService, data_loader.DeviceShardLoader):
        assert not getattr(cls, "__abstractmethods__", None), cls


def test_contract_method_names_and_arguments_match_upstream():
    want = {
        interfaces.AggregationServiceInterface: {"aggregate_updates": ["updates", "weights"], "validate_update": ["update"],
                                                 "compress_global_model": ["model"],
                                                 "calculate_convergence_metrics": ["old_model", "new_model"]},
        interfaces.ModelInterface: {"get_model_weights": [], "set_model_weights": ["weights"], "get_parameter_count": [],
                                    "estimate_memory_usage": []},
        interfaces.DataLoaderInterface: {"load_training_data": ["client_id"], "load_validation_data": [],
                                         "get_data_statistics": ["client_id"]},
        interfaces.PrivacyEngineInterface: {"add_noise": ["gradients", "epsilon", "delta"], "clip_gradients": ["gradients", "max_norm"],
                                            "calculate_privacy_budget": ["epsilon", "delta", "steps"],
                                            "validate_privacy_parameters": ["epsilon", "delta"]},
        interfaces.CompressionInterface: {"compress_weights": ["weights"], "decompress_weights": ["compressed_data"],
                                          "get_compression_ratio": ["original_size", "compressed_size"]},
    }
    for cls, methods in want.items():
        assert set(cls.__abstractmethods__) == set(methods), cls
        for name, args in methods.items():
            assert list(inspect.signature(getattr(cls, name)).parameters)[1:] == args, (cls, name)


def test_noise_seeds_default_to_fresh_entropy():
    """a constant default seed makes the DP noise reproducible by anyone.  Construction needs no GPU."""
    import torch
    a = privacy.GaussianNoiseGenerator(torch.device("cuda"))
    b = privacy.GaussianNoiseGenerator(torch.device("cuda"))
    assert a.seed != b.seed and a.seed not in (0, 42)
    assert privacy.GaussianNoiseGenerator(torch.device("cuda"), seed=7).seed == 7
    sig = inspect.signature(privacy.create_privacy_engine)
    assert sig.parameters["seed"].default is None
    assert inspect.signature(privacy.DifferentialPrivacyEngine.__init__).parameters["seed"].default is None
