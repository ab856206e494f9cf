"""Data side of the hot path: the reference's ``DataPartitioner``
(src/shared/data_loader.py:65-264: iid / Dirichlet non-iid / pathological splits) and the per-batch
ToTensor + Normalize + host->device copy of its loaders (:298-306, :454-464, training.py:186), restated as

  * ``DataPartitioner`` -- same constructor and algorithms, consuming ``random`` / ``numpy.random`` in the same order, so
    that the same seeds give the same client index lists as upstream (pinned by tests/golden/partition.npz);
  * ``DeviceShardBuilder`` -- the raw uint8 dataset is uploaded ONCE and a single gather + normalise kernel
    (``csrc/shards.cu``) writes the packed, device-resident sample store that ``FederatedRoundEngine.attach_device_shards``
    hands to the training kernels: no per-batch transform, no per-batch copy.

  * ``DeviceShardLoader`` -- the reference's ``MNISTDataLoader`` / ``CIFAR10DataLoader`` contract (``DataLoaderInterface``:
    load_training_data / load_validation_data / get_data_statistics, :267-420) over arrays the caller already holds.

Dataset download (torchvision, network) and the CIFAR train-time RandomCrop / RandomHorizontalFlip augmentation are not
part of this module."""
from __future__ import annotations

import logging
import random
from collections import defaultdict
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from .interfaces import DataLoaderInterface

logger = logging.getLogger(__name__)

MNIST_MEAN_STD = ((0.1307,), (0.3081,))                                        # data_loader.py:298-301
CIFAR10_MEAN_STD = ((0.4914, 0.4822, 0.4465), (0.2023, 0.1994, 0.2010))        # data_loader.py:454-464


class FederatedDataset(torch.utils.data.Dataset):
    """One client's view of the base dataset (src/shared/data_loader.py:22-62): an index list, no copy."""

    def __init__(self, base_dataset, client_id: str, indices: List[int]):
        self.base_dataset, self.client_id, self.indices = base_dataset, client_id, indices

    def __len__(self) -> int:
        return len(self.indices)

    def __getitem__(self, idx: int):
        return self.base_dataset[self.indices[idx]]

    def get_statistics(self) -> Dict[str, Any]:
        counts: Dict[int, int] = defaultdict(int)
        for i in self.indices:
            counts[int(self.base_dataset[i][1])] += 1
        return {"client_id": self.client_id, "total_samples": len(self.indices), "class_distribution": dict(counts),
                "num_classes": len(counts)}


class DataPartitioner:
    """src/shared/data_loader.py:65-264.  ``dataset`` is anything indexable as ``dataset[i] -> (x, label)``; pass
    ``labels=`` to skip the per-item label extraction (:101-107) for large datasets."""

    def __init__(self, dataset, num_clients: int, partition_strategy: str = "iid", alpha: float = 0.5,
                 min_samples_per_client: int = 10, labels: Optional[Sequence[int]] = None):
        self.dataset = dataset
        self.num_clients = num_clients
        self.partition_strategy = partition_strategy
        self.alpha = alpha
        self.min_samples_per_client = min_samples_per_client
        self.labels = [int(v) for v in labels] if labels is not None else self._extract_labels()
        self.num_classes = len(set(self.labels))
        self.client_indices = self._create_partitions()
        logger.info(f"Created {partition_strategy} partitions for {num_clients} clients")

    def _len(self) -> int:
        return len(self.labels)

    def _extract_labels(self) -> List[int]:
        return [int(self.dataset[i][1]) for i in range(len(self.dataset))]

    def _create_partitions(self) -> Dict[int, List[int]]:
        if self.partition_strategy == "iid":
            return self._create_iid_partitions()
        if self.partition_strategy == "non_iid":
            return self._create_non_iid_partitions()
        if self.partition_strategy == "pathological":
            return self._create_pathological_partitions()
        raise ValueError(f"Unknown partition strategy: {self.partition_strategy}")

    def _create_iid_partitions(self) -> Dict[int, List[int]]:
        indices = list(range(self._len()))
        random.shuffle(indices)                                                  # :121
        per = len(indices) // self.num_clients
        out = {}
        for cid in range(self.num_clients):
            start = cid * per
            end = len(indices) if cid == self.num_clients - 1 else start + per   # last client takes the remainder (:128-132)
            out[cid] = indices[start:end]
        return out

    def _create_non_iid_partitions(self) -> Dict[int, List[int]]:
        class_indices = defaultdict(list)
        for idx, label in enumerate(self.labels):
            class_indices[label].append(idx)
        client_indices = defaultdict(list)
        for _, indices in class_indices.items():                                 # insertion order of first appearance (:142-145)
            proportions = np.random.dirichlet([self.alpha] * self.num_clients)   # :151
            proportions = np.maximum(proportions, self.min_samples_per_client / len(indices))
            proportions = proportions / proportions.sum()
            np.random.shuffle(indices)                                           # :159
            start = 0
            for cid in range(self.num_clients):
                n = int(proportions[cid] * len(indices))
                end = len(indices) if cid == self.num_clients - 1 else start + n
                client_indices[cid].extend(indices[start:end])
                start = end
        for cid in client_indices:
            random.shuffle(client_indices[cid])                                  # :174-175
        return dict(client_indices)

    def _create_pathological_partitions(self) -> Dict[int, List[int]]:
        class_indices = defaultdict(list)
        for idx, label in enumerate(self.labels):
            class_indices[label].append(idx)
        client_indices = defaultdict(list)
        classes_per_client = max(1, self.num_classes // self.num_clients)
        class_list = list(class_indices.keys())
        random.shuffle(class_list)                                               # :192
        assignments = {}
        for cid in range(self.num_clients):
            start_class = (cid * classes_per_client) % self.num_classes
            assignments[cid] = [class_list[(start_class + i) % self.num_classes] for i in range(classes_per_client)]
        for cid, assigned in assignments.items():
            for label in assigned:
                indices = class_indices[label].copy()
                random.shuffle(indices)                                          # :208
                holders = sum(1 for classes in assignments.values() if label in classes)
                client_indices[cid].extend(indices[:len(indices) // holders])
        for cid in range(self.num_clients):
            if len(client_indices[cid]) < self.min_samples_per_client:           # :220-239
                used = set()
                for v in client_indices.values():
                    used.update(v)
                available = list(set(range(self._len())) - used)
                need = self.min_samples_per_client - len(client_indices[cid])
                if available:
                    client_indices[cid].extend(random.sample(available, min(need, len(available))))
        return dict(client_indices)

    def get_client_dataset(self, client_id: int) -> FederatedDataset:
        if client_id not in self.client_indices:
            raise ValueError(f"Client {client_id} not found")
        return FederatedDataset(self.dataset, str(client_id), self.client_indices[client_id])

    def get_partition_statistics(self) -> Dict[str, Any]:
        stats = {"num_clients": self.num_clients, "partition_strategy": self.partition_strategy,
                 "total_samples": self._len(), "num_classes": self.num_classes, "client_statistics": {}}
        for cid, indices in self.client_indices.items():
            counts: Dict[int, int] = defaultdict(int)
            for i in indices:
                counts[self.labels[i]] += 1
            stats["client_statistics"][cid] = {"client_id": str(cid), "total_samples": len(indices),
                                               "class_distribution": dict(counts), "num_classes": len(counts)}
        return stats


class DeviceShardBuilder:
    """Raw uint8 dataset resident on the device + one gather/normalise launch per shard set.

    ``images``: uint8 ``[N, H, W]`` (MNIST ``.data``), ``[N, H, W, C]`` (CIFAR10 ``.data``) or ``[N, C, H, W]`` with
    ``channels_last=False``; ``labels``: ``[N]`` integers; ``mean`` / ``std``: per channel (``MNIST_MEAN_STD`` ...)."""

    def __init__(self, images, labels, mean: Sequence[float], std: Sequence[float], device=None, channels_last: bool = True):
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        L.ensure_device(dev)
        img = torch.as_tensor(np.asarray(images) if not isinstance(images, torch.Tensor) else images)
        if img.dtype != torch.uint8:
            raise L.FlbError("DeviceShardBuilder: images must be uint8 (the raw dataset); normalisation happens on the device")
        if img.dim() == 3:
            img = img.unsqueeze(-1) if channels_last else img.unsqueeze(1)
        if img.dim() != 4:
            raise L.FlbError("DeviceShardBuilder: images must be [N, H, W], [N, H, W, C] or [N, C, H, W]")
        self.hwc = bool(channels_last)
        self.N = int(img.shape[0])
        self.H, self.W, self.C = (int(img.shape[1]), int(img.shape[2]), int(img.shape[3])) if self.hwc else \
            (int(img.shape[2]), int(img.shape[3]), int(img.shape[1]))
        if len(mean) != self.C or len(std) != self.C:
            raise L.FlbError(f"DeviceShardBuilder: need {self.C} mean / std values")
        self.device = dev
        self.raw = img.contiguous().to(dev)
        self.labels = torch.as_tensor(np.asarray(labels) if not isinstance(labels, torch.Tensor) else labels).to(torch.int64).to(dev)
        self.mean = torch.tensor([float(v) for v in mean], dtype=torch.float32, device=dev)
        self.std = torch.tensor([float(v) for v in std], dtype=torch.float32, device=dev)

    def build(self, client_indices: Dict[int, List[int]], client_ids: Sequence[int], validation_split: float = 0.0,
              seed: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor, List[int]]:
        """Packed training store for ``client_ids`` (in that order): ``x [sum N_c, C*H*W]`` fp32 normalised, ``y [sum N_c]``
        int32, and the per-client counts.  ``validation_split`` holds out the reference's ``int(n * split)`` samples per
        client (data_loader.py:344-352; a seeded permutation stands in for ``random_split``)."""
        gen = torch.Generator().manual_seed(seed) if seed is not None else None
        chunks, sizes = [], []
        for cid in client_ids:
            idx = torch.tensor(client_indices[cid], dtype=torch.int64)
            if validation_split > 0:
                n_val = int(idx.numel() * validation_split)
                perm = torch.randperm(idx.numel(), generator=gen)
                idx = idx[perm[:idx.numel() - n_val]]
            chunks.append(idx)
            sizes.append(int(idx.numel()))
        all_idx = torch.cat(chunks) if chunks else torch.zeros(0, dtype=torch.int64)
        if all_idx.numel() and (int(all_idx.min()) < 0 or int(all_idx.max()) >= self.N):
            raise L.FlbError("DeviceShardBuilder.build: sample index out of range")
        M = int(all_idx.numel())
        idx_dev = all_idx.to(self.device)
        x = torch.empty((max(M, 1), self.C * self.H * self.W), dtype=torch.float32, device=self.device)
        y = torch.empty(max(M, 1), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            st = L.stream_ptr(self.device)
            L.call("flb_gather_normalize_u8", L.ptr(self.raw), self.N, self.H, self.W, self.C, int(self.hwc), L.ptr(idx_dev), M,
                   L.ptr(self.mean), L.ptr(self.std), L.ptr(x), st)
            L.call("flb_gather_labels", L.ptr(self.labels), L.ptr(idx_dev), L.ptr(y), M, st)
        return x, y, sizes


class DeviceBatches:
    """What ``DataLoader(dataset, batch_size, shuffle)`` yields (data_loader.py:356-362), over tensors that already sit
    normalised on the device: every pass re-draws the permutation when ``shuffle`` is set; batches are slices, no copies
    to the host and back.  ``LocalTrainer`` accepts it like any loader."""

    def __init__(self, x: torch.Tensor, y: torch.Tensor, batch_size: int, shuffle: bool):
        self.x, self.y, self.batch_size, self.shuffle = x, y, int(batch_size), shuffle
        self.dataset = y                    # len(loader.dataset), as callers of a DataLoader use it

    def __len__(self) -> int:
        return (int(self.y.shape[0]) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = int(self.y.shape[0])
        x, y = self.x, self.y
        if self.shuffle and n:
            perm = torch.randperm(n, device=x.device)
            x, y = x[perm], y[perm]
        for i in range(0, n, self.batch_size):
            yield x[i:i + self.batch_size], y[i:i + self.batch_size]


class DeviceShardLoader(DataLoaderInterface):
    """``MNISTDataLoader`` / ``CIFAR10DataLoader`` (data_loader.py:267-420, :423-600) without the download: the caller
    passes the raw uint8 training (and optionally test) arrays; partitioning follows ``DataPartitioner`` and the
    per-client tensors are built by ONE gather + normalise launch each (``DeviceShardBuilder``) and cached on the device.
    Client ids are parsed like upstream (:333): ``"client-3"`` or ``"3"``."""

    def __init__(self, images, labels, mean: Sequence[float], std: Sequence[float], num_clients: int = 10,
                 partition_strategy: str = "iid", batch_size: int = 32, validation_split: float = 0.1, device=None,
                 test_images=None, test_labels=None, channels_last: bool = True):
        self.num_clients, self.partition_strategy = num_clients, partition_strategy
        self.batch_size, self.validation_split = batch_size, validation_split
        self.builder = DeviceShardBuilder(images, labels, mean, std, device, channels_last)
        self.test_builder = (DeviceShardBuilder(test_images, test_labels, mean, std, device, channels_last)
                             if test_images is not None else None)
        self.partitioner = DataPartitioner(None, num_clients, partition_strategy,
                                           labels=[int(v) for v in np.asarray(labels).reshape(-1)])
        self._splits: Dict[int, Tuple[List[int], List[int]]] = {}
        self._shape = (self.builder.C, self.builder.H, self.builder.W)

    @staticmethod
    def _client_index(client_id) -> int:
        cid = str(client_id)
        return int(cid.split("-")[-1]) if "-" in cid else int(cid)

    def _split(self, idx: int) -> Tuple[List[int], List[int]]:
        """(train, validation) index lists of one client: ``int(n * split)`` held out (:344-352), drawn once per client
        so that the two loaders never overlap (upstream re-draws ``random_split`` on every call)."""
        if idx >= self.num_clients:
            raise ValueError(f"Client ID {idx} exceeds number of clients {self.num_clients}")
        if idx not in self._splits:
            ind = list(self.partitioner.client_indices[idx])
            n_val = int(len(ind) * self.validation_split) if self.validation_split > 0 else min(100, len(ind) // 10)
            perm = torch.randperm(len(ind)).tolist() if self.validation_split > 0 else list(range(len(ind)))
            val = [ind[i] for i in perm[:n_val]]
            train = [ind[i] for i in perm[n_val:]] if self.validation_split > 0 else ind
            self._splits[idx] = (train, val)
        return self._splits[idx]

    def _batches(self, builder: DeviceShardBuilder, indices: List[int], shuffle: bool) -> DeviceBatches:
        x, y, sizes = builder.build({0: indices}, [0])
        n = sizes[0]
        return DeviceBatches(x[:n].view((n,) + self._shape), y[:n].to(torch.int64), self.batch_size, shuffle)

    def load_training_data(self, client_id) -> DeviceBatches:
        return self._batches(self.builder, self._split(self._client_index(client_id))[0], shuffle=True)

    def load_validation_data(self, client_id=None) -> DeviceBatches:
        if client_id:
            return self._batches(self.builder, self._split(self._client_index(client_id))[1], shuffle=False)
        if self.test_builder is None:
            raise ValueError("no test set was given: pass test_images / test_labels for the global validation loader")
        return self._batches(self.test_builder, list(range(self.test_builder.N)), shuffle=False)

    def get_data_statistics(self, client_id) -> Dict[str, Any]:
        try:
            idx = self._client_index(client_id)
            counts: Dict[int, int] = defaultdict(int)
            for i in self.partitioner.client_indices[idx]:
                counts[self.partitioner.labels[i]] += 1
            return {"client_id": str(idx), "total_samples": len(self.partitioner.client_indices[idx]),
                    "class_distribution": dict(counts), "num_classes": len(counts)}
        except Exception as e:
            logger.error(f"Failed to get data statistics for client {client_id}: {e}")
            return {}
