"""Update compression behind the reference's ``ModelCompressionService`` surface (``src/shared/compression.py``:
BaseCompressor :19-59, QuantizationCompressor :123-247, TopKSparsificationCompressor :250-368,
ModelCompressionService :371-470, create_compression_service :473-485).

The arithmetic (per-tensor affine quantisation, exact top-k by |x|, and their inverses) runs in the CUDA kernels of
``csrc/codec.cu`` / ``csrc/topk.cu`` on the flattened update, all layers in one launch; the byte envelope (a pickle of
``{compressed_data, metadata}`` with the reference's metadata keys) is host code and stays compatible: bytes produced here
decompress with the reference's classes and vice versa for the quantisation / top-k algorithms.

Differences, on purpose: weights must live on a CUDA device (no CPU fallback); quantisation supports 1..8 bits (the codes are
uint8, as upstream stores them for ``bits <= 8``); top-k returns its (value, index) pairs in index order rather than by
magnitude (the dense reconstruction -- all the reference ever consumes -- is identical; torch.topk's tie order is
unspecified anyway); the LZ4 codec (a byte-level wire codec around ``torch.save``, never invoked by the client path --
``compression_ratio=0.8`` placeholder at ``src/client/federated_trainer.py:484``) is out of scope and raises."""
from __future__ import annotations

import io
import logging
import pickle
from abc import ABC, abstractmethod
from collections import OrderedDict
from typing import Any, Dict, Tuple

import torch

from . import _lib as L
from . import ops
from .interfaces import CompressionInterface
from .layout import ParamLayout
from .models import ModelWeights

logger = logging.getLogger(__name__)


class CompressionError(Exception):
    pass


class BaseCompressor(ABC):
    @abstractmethod
    def compress(self, weights: ModelWeights) -> Tuple[bytes, Dict[str, Any]]:
        ...

    @abstractmethod
    def decompress(self, compressed_data: bytes, metadata: Dict[str, Any]) -> ModelWeights:
        ...

    @abstractmethod
    def get_compression_name(self) -> str:
        ...


def _pack(weights: ModelWeights):
    """dict of CUDA tensors -> (layout, [1, ld] row, device)."""
    if not weights:
        raise ValueError("empty weights")
    first = next(iter(weights.values()))
    if not first.is_cuda:
        raise L.FlbError("compression kernels need CUDA tensors (no CPU fallback)")
    layout = ParamLayout(OrderedDict((k, tuple(v.shape)) for k, v in weights.items()))
    row = torch.zeros((1, layout.ld), dtype=torch.float32, device=first.device)
    layout.flatten_into(row[0], {k: v.to(torch.float32) for k, v in weights.items()})
    return layout, row, first.device


class QuantizationCompressor(BaseCompressor):
    def __init__(self, bits: int = 8, symmetric: bool = True):
        self.bits = max(1, min(32, bits))
        self.symmetric = symmetric
        self.levels = 2 ** self.bits

    def compress(self, weights: ModelWeights) -> Tuple[bytes, Dict[str, Any]]:
        try:
            if self.bits > 8:
                raise L.FlbError("quantisation kernels store uint8 codes: bits must be in 1..8")
            layout, row, dev = _pack(weights)
            q, scale, zp = ops.q8_quantize(row, layout.seg_off(dev), P=layout.P, bits=self.bits, symmetric=self.symmetric)
            scale_h, zp_h = scale[0].cpu().tolist(), zp[0].cpu().tolist()          # the envelope's one device -> host read
            q_host = q[0, :layout.P].cpu()
            compressed, params = {}, {}
            for i, name in enumerate(layout.names):
                o, shp = layout.offsets[name], layout.shapes[name]
                n = 1
                for s_ in shp:
                    n *= s_
                compressed[name] = q_host[o:o + n].reshape(shp).clone()
                params[name] = {"scale": float(scale_h[i]), "zero_point": int(zp_h[i]), "original_shape": torch.Size(shp),
                                "original_dtype": str(weights[name].dtype)}
            buf = io.BytesIO()
            torch.save(compressed, buf)
            return buf.getvalue(), {"algorithm": self.get_compression_name(), "bits": self.bits, "symmetric": self.symmetric,
                                    "quantization_params": params}
        except Exception as e:
            logger.error(f"Quantization compression failed: {str(e)}")
            raise CompressionError(f"Quantization compression failed: {str(e)}")

    def decompress(self, compressed_data: bytes, metadata: Dict[str, Any], device=None) -> ModelWeights:
        try:
            codes = torch.load(io.BytesIO(compressed_data), map_location="cpu")
            params = metadata["quantization_params"]
            dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
            layout = ParamLayout(OrderedDict((k, tuple(params[k]["original_shape"])) for k in codes))
            q = torch.zeros((1, layout.ld), dtype=torch.uint8)
            for name in layout.names:
                o = layout.offsets[name]
                flat = codes[name].reshape(-1)
                if flat.dtype != torch.uint8:
                    raise L.FlbError("only uint8 codes (bits <= 8) are supported by the kernels")
                q[0, o:o + flat.numel()] = flat
            q = q.to(dev)
            scale = torch.tensor([[float(params[k]["scale"]) for k in layout.names]], dtype=torch.float32, device=dev)
            zp = torch.tensor([[float(params[k]["zero_point"]) for k in layout.names]], dtype=torch.float32, device=dev)
            dense = ops.q8_dequantize(q, scale, zp, layout.seg_off(dev), P=layout.P)
            out = layout.unflatten(dense[0])
            for name in out:
                if "float64" in params[name]["original_dtype"]:
                    out[name] = out[name].to(torch.float64)
            return out
        except Exception as e:
            logger.error(f"Quantization decompression failed: {str(e)}")
            raise CompressionError(f"Quantization decompression failed: {str(e)}")

    def get_compression_name(self) -> str:
        return f"quantization_{self.bits}bit"


class TopKSparsificationCompressor(BaseCompressor):
    def __init__(self, sparsity_ratio: float = 0.9):
        self.sparsity_ratio = max(0.0, min(1.0, sparsity_ratio))

    def compress(self, weights: ModelWeights) -> Tuple[bytes, Dict[str, Any]]:
        try:
            layout, row, dev = _pack(weights)
            counts = []
            for name in layout.names:
                n = 1
                for s_ in layout.shapes[name]:
                    n *= s_
                counts.append(min(n, ops.topk_count(n, self.sparsity_ratio)))
            idx, val, off_t, _ = ops.topk_select(row, layout.seg_off(dev), counts, P=layout.P)
            idx_h, val_h, offs = idx[0].cpu(), val[0].cpu(), off_t.cpu().tolist()
            compressed, params = {}, {}
            for i, name in enumerate(layout.names):
                a, b = offs[i], offs[i + 1]
                compressed[name] = {"values": val_h[a:b].clone(), "indices": idx_h[a:b].to(torch.int64)}
                params[name] = {"original_shape": torch.Size(layout.shapes[name]), "original_dtype": str(weights[name].dtype),
                                "sparsity_ratio": self.sparsity_ratio}
            buf = io.BytesIO()
            pickle.dump(compressed, buf, protocol=pickle.HIGHEST_PROTOCOL)
            return buf.getvalue(), {"algorithm": self.get_compression_name(), "sparsity_ratio": self.sparsity_ratio,
                                    "sparsification_params": params}
        except Exception as e:
            logger.error(f"Top-K sparsification failed: {str(e)}")
            raise CompressionError(f"Top-K sparsification failed: {str(e)}")

    def decompress(self, compressed_data: bytes, metadata: Dict[str, Any], device=None) -> ModelWeights:
        try:
            sparse = pickle.load(io.BytesIO(compressed_data))
            params = metadata["sparsification_params"]
            dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
            layout = ParamLayout(OrderedDict((k, tuple(params[k]["original_shape"])) for k in sparse))
            counts = [int(sparse[k]["indices"].numel()) for k in layout.names]
            offs = [0]
            for c in counts:
                offs.append(offs[-1] + c)
            ldk = max(offs[-1], 1)
            idx = torch.zeros((1, ldk), dtype=torch.int32)
            val = torch.zeros((1, ldk), dtype=torch.float32)
            for i, name in enumerate(layout.names):
                idx[0, offs[i]:offs[i + 1]] = sparse[name]["indices"].to(torch.int32)
                val[0, offs[i]:offs[i + 1]] = sparse[name]["values"].to(torch.float32)
            kk_t = torch.tensor(counts, dtype=torch.int32, device=dev)
            off_t = torch.tensor(offs, dtype=torch.int64, device=dev)
            dense = ops.topk_scatter(idx.to(dev), val.to(dev), layout.seg_off(dev), kk_t, off_t, layout.P, layout.ld)
            out = layout.unflatten(dense[0])
            for name in out:
                if "float64" in params[name]["original_dtype"]:
                    out[name] = out[name].to(torch.float64)
            return out
        except Exception as e:
            logger.error(f"Top-K desparsification failed: {str(e)}")
            raise CompressionError(f"Top-K desparsification failed: {str(e)}")

    def get_compression_name(self) -> str:
        return f"topk_sparsification_{self.sparsity_ratio}"


class ModelCompressionService(CompressionInterface):
    """src/shared/compression.py:371-470 (same methods and envelope)."""

    def __init__(self, algorithm: str = "quantization", **kwargs):
        self.algorithm = algorithm
        self.compressor = self._create_compressor(algorithm, **kwargs)

    def _create_compressor(self, algorithm: str, **kwargs) -> BaseCompressor:
        if algorithm == "quantization":
            return QuantizationCompressor(**kwargs)
        if algorithm == "topk":
            return TopKSparsificationCompressor(**kwargs)
        if algorithm == "lz4":
            raise ValueError("the lz4 byte codec is outside the B200 hot path (see DESIGN.md, out of scope); "
                             "use 'quantization' or 'topk'")
        raise ValueError(f"Unknown compression algorithm: {algorithm}")

    def compress_weights(self, weights: ModelWeights) -> bytes:
        try:
            data, metadata = self.compressor.compress(weights)
            buf = io.BytesIO()
            pickle.dump({"compressed_data": data, "metadata": metadata}, buf, protocol=pickle.HIGHEST_PROTOCOL)
            return buf.getvalue()
        except Exception as e:
            logger.error(f"Weight compression failed: {str(e)}")
            raise CompressionError(f"Weight compression failed: {str(e)}")

    def decompress_weights(self, compressed_data: bytes) -> ModelWeights:
        try:
            package = pickle.load(io.BytesIO(compressed_data))
            data, metadata = package["compressed_data"], package["metadata"]
            algorithm = metadata["algorithm"]
            compressor = self.compressor
            if algorithm != self.compressor.get_compression_name():
                if algorithm.startswith("quantization"):
                    compressor = QuantizationCompressor(bits=metadata["bits"], symmetric=metadata["symmetric"])
                elif algorithm.startswith("topk"):
                    compressor = TopKSparsificationCompressor(sparsity_ratio=metadata["sparsity_ratio"])
            return compressor.decompress(data, metadata)
        except Exception as e:
            logger.error(f"Weight decompression failed: {str(e)}")
            raise CompressionError(f"Weight decompression failed: {str(e)}")

    def get_compression_ratio(self, original_size: int, compressed_size: int) -> float:
        if original_size == 0:
            return 0.0
        return compressed_size / original_size

    def estimate_compression_ratio(self, weights: ModelWeights) -> float:
        try:
            original = sum(t.numel() * t.element_size() for t in weights.values())
            return self.get_compression_ratio(original, len(self.compress_weights(weights)))
        except Exception as e:
            logger.error(f"Compression ratio estimation failed: {str(e)}")
            return 1.0


def create_compression_service(algorithm: str = "quantization", **kwargs) -> ModelCompressionService:
    return ModelCompressionService(algorithm, **kwargs)
