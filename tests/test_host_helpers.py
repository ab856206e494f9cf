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


def test_peer_region_layout():
    """flb_p2p_region_layout (include/flb.h): the per-rank peer-memory region of the fused multi-GPU FedAvg -- flags for every
    (chunk, source rank), one flag per chunk of the global row, a G-row inbox and the global row, 256-byte aligned and disjoint."""
    import ctypes as C
    from flb200 import _lib as L
    lib = L.load()
    for ld, world, chunk in ((421664, 2, 2048), (1470912, 8, 2048), (100_000_000, 8, 98304), (32, 16, 4)):
        lay = L.P2pLayout()
        nbytes = lib.flb_p2p_region_layout(ld, world, chunk, C.byref(lay))
        nchunks = -(-ld // chunk)
        assert nbytes == lay.bytes and (lay.ld, lay.world, lay.chunk) == (ld, world, chunk)
        assert lay.off_flags_a == 0 and lay.off_flags_b >= 4 * nchunks * world
        assert lay.off_inbox >= lay.off_flags_b + 4 * nchunks and lay.off_global >= lay.off_inbox + 4 * world * ld
        assert lay.bytes >= lay.off_global + 4 * ld
        assert all(o % 256 == 0 for o in (lay.off_flags_b, lay.off_inbox, lay.off_global, lay.bytes))
    bad = L.P2pLayout()
    for args in ((30, 2, 2048), (32, 0, 2048), (32, 17, 2048), (32, 2, 6)):          # ld % 4, world range, chunk % 4
        assert lib.flb_p2p_region_layout(*args, C.byref(bad)) == -1


def test_round_engine_falls_back_to_nccl_path_without_peer_memory(monkeypatch):
    """PeerFedAvg refuses a non-CUDA device before any collective is issued, on every rank alike, so the engine's constructor
    cannot dead-lock half-way into the handle exchange."""
    import pytest
    from flb200.p2p import PeerFedAvg
    with pytest.raises(L_FlbError()):
        PeerFedAvg(64, torch.device("cpu"), 0, 2, None)
    with pytest.raises(L_FlbError()):
        PeerFedAvg(64, torch.device("cpu"), 0, 1, None)


def L_FlbError():
    from flb200 import FlbError
    return FlbError
