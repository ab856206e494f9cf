"""Data side: the partitioner against golden client index lists produced by the unmodified reference
DataPartitioner (CPU), and the device shard builder against the ToTensor + Normalize formula (GPU)."""
import random

import numpy as np
import pytest
import torch

from conftest import load_golden

import flb200  # noqa: F401


@pytest.mark.parametrize("strategy", ["iid", "non_iid", "pathological"])
@pytest.mark.parametrize("num_clients", [7, 20])
def test_partitioner_matches_reference_golden(strategy, num_clients):
    from flb200.data_loader import DataPartitioner
    gold = load_golden("partition.npz")
    labels = gold["labels"]
    random.seed(5)
    np.random.seed(6)
    part = DataPartitioner(None, num_clients, strategy, labels=labels)
    assert sorted(part.client_indices) == list(range(num_clients))
    for cid, idx in part.client_indices.items():
        np.testing.assert_array_equal(np.asarray(idx, dtype=np.int64), gold[f"{strategy}/{num_clients}/{cid}"])
    stats = part.get_partition_statistics()
    assert stats["num_clients"] == num_clients and stats["total_samples"] == 3000
    assert stats["client_statistics"][0]["total_samples"] == len(part.client_indices[0])
    with pytest.raises(ValueError, match="Unknown partition strategy"):
        DataPartitioner(None, 3, "sorted", labels=labels)


def test_client_dataset_view():
    from flb200.data_loader import DataPartitioner, FederatedDataset
    y = [i % 5 for i in range(200)]
    base = [(torch.full((1,), float(i)), y[i]) for i in range(200)]
    random.seed(1)
    part = DataPartitioner(base, 4, "iid")
    ds = part.get_client_dataset(2)
    assert isinstance(ds, FederatedDataset) and ds.client_id == "2" and len(ds) == len(part.client_indices[2])
    x0, y0 = ds[0]
    assert int(x0.item()) == part.client_indices[2][0] and y0 == y[part.client_indices[2][0]]
    st = ds.get_statistics()
    assert st["total_samples"] == len(ds) and sum(st["class_distribution"].values()) == len(ds)
    assert st == part.get_partition_statistics()["client_statistics"][2]
    with pytest.raises(ValueError, match="Client 9 not found"):
        part.get_client_dataset(9)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["mnist", "cifar"])
def test_device_shard_builder_is_totensor_normalize(cuda_device, kind):
    from flb200.data_loader import CIFAR10_MEAN_STD, MNIST_MEAN_STD, DataPartitioner, DeviceShardBuilder
    g = torch.Generator().manual_seed(1)
    N = 500
    if kind == "mnist":
        raw = torch.randint(0, 256, (N, 28, 28), generator=g, dtype=torch.uint8)
        mean, std = MNIST_MEAN_STD
        chw = raw.unsqueeze(1)
    else:
        raw = torch.randint(0, 256, (N, 32, 32, 3), generator=g, dtype=torch.uint8)
        mean, std = CIFAR10_MEAN_STD
        chw = raw.permute(0, 3, 1, 2)
    labels = torch.randint(0, 10, (N,), generator=g)
    random.seed(2)
    np.random.seed(3)
    part = DataPartitioner(None, 4, "non_iid", labels=labels.tolist())
    sb = DeviceShardBuilder(raw, labels, mean, std, cuda_device)
    ids = [2, 0, 3]
    x, y, sizes = sb.build(part.client_indices, ids)
    assert sizes == [len(part.client_indices[c]) for c in ids]
    idx = torch.tensor(sum((part.client_indices[c] for c in ids), []))
    # torchvision: ToTensor = uint8 -> float / 255; Normalize = (t - mean) / std, each in fp32
    t = chw[idx].to(torch.float32).div(255)
    ref = (t - torch.tensor(mean).view(1, -1, 1, 1)) / torch.tensor(std).view(1, -1, 1, 1)
    assert torch.equal(x[:idx.numel()].cpu(), ref.reshape(idx.numel(), -1))          # bit-exact
    assert torch.equal(y[:idx.numel()].cpu().long(), labels[idx])
    # validation hold-out and attaching the store to an engine without a copy
    x2, y2, sizes2 = sb.build(part.client_indices, ids, validation_split=0.1, seed=7)
    assert sizes2 == [n - int(n * 0.1) for n in sizes]
    from flb200.simulation import FederatedRoundEngine
    model = "simple_cnn" if kind == "mnist" else "cifar10_cnn"
    eng = FederatedRoundEngine(model, 3, cuda_device, dp_mode="none", dropout_rate=0.0, precision="fp32", batch_size=16)
    from flb200.models_pytorch import ModelFactory
    torch.manual_seed(0)
    eng.set_global_weights(ModelFactory.create_model(model).get_model_weights())
    eng.attach_device_shards(x2, y2, sizes2)
    assert eng.trainer.x.data_ptr() == x2.data_ptr()
    out = eng.run_round()
    assert out["samples"] == sizes2 and all(np.isfinite(out["losses"]))
