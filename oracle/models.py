"""Oracle (test infrastructure): functional restatement of the two CNNs on the hot path.

Follows ``src/shared/models_pytorch.py``:
  * SimpleCNN   layers ``:69-80``, forward ``:82-97``
  * CIFAR10CNN  layers ``:110-134``, forward ``:136-165``
Parameter order is the reference's ``named_parameters()`` registration order
(``get_model_weights`` ``:25-27``); BatchNorm running statistics are buffers and
are NOT part of the federated weights (``:25-27`` iterates parameters only).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

# name -> shape, in reference registration order -----------------------------------------


def simple_cnn_spec(num_classes: int = 10) -> "OrderedDict[str, Tuple[int, ...]]":
    # models_pytorch.py:69-80
    return OrderedDict([
        ("conv1.weight", (32, 1, 3, 3)), ("conv1.bias", (32,)),
        ("conv2.weight", (64, 32, 3, 3)), ("conv2.bias", (64,)),
        ("fc1.weight", (128, 64 * 7 * 7)), ("fc1.bias", (128,)),
        ("fc2.weight", (num_classes, 128)), ("fc2.bias", (num_classes,)),
    ])


_CIFAR_CONVS = [(3, 32), (32, 32), (32, 64), (64, 64), (64, 128), (128, 128)]


def cifar10_cnn_spec(num_classes: int = 10) -> "OrderedDict[str, Tuple[int, ...]]":
    # models_pytorch.py:110-134 (conv_i followed by bn_i, then the three linears)
    spec: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    for i, (cin, cout) in enumerate(_CIFAR_CONVS, start=1):
        spec[f"conv{i}.weight"] = (cout, cin, 3, 3)
        spec[f"conv{i}.bias"] = (cout,)
        spec[f"bn{i}.weight"] = (cout,)
        spec[f"bn{i}.bias"] = (cout,)
    spec["fc1.weight"] = (512, 128 * 4 * 4)
    spec["fc1.bias"] = (512,)
    spec["fc2.weight"] = (256, 512)
    spec["fc2.bias"] = (256,)
    spec["fc3.weight"] = (num_classes, 256)
    spec["fc3.bias"] = (num_classes,)
    return spec


def model_spec(name: str, num_classes: int = 10):
    if name == "simple_cnn":
        return simple_cnn_spec(num_classes)
    if name == "cifar10_cnn":
        return cifar10_cnn_spec(num_classes)
    raise ValueError(f"Unknown model: {name}")


def input_shape(name: str) -> Tuple[int, int, int]:
    return (1, 28, 28) if name == "simple_cnn" else (3, 32, 32)


def init_weights(name: str, seed: int = 0, num_classes: int = 10) -> Dict[str, torch.Tensor]:
    """torch's default init for Conv2d / Linear (kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), +)),
    BatchNorm weight 1 / bias 0.  Not bit-identical to constructing the reference module
    (different RNG consumption order is possible) -- golden tests ship explicit weights."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    spec = model_spec(name, num_classes)
    for pname, shape in spec.items():
        layer, kind = pname.split(".")
        if layer.startswith("bn"):
            out[pname] = torch.ones(shape) if kind == "weight" else torch.zeros(shape)
            continue
        wshape = spec[f"{layer}.weight"]
        fan_in = int(math.prod(wshape[1:]))
        bound = 1.0 / math.sqrt(fan_in)
        out[pname] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return out


def new_bn_state(name: str) -> Dict[str, torch.Tensor]:
    """Client-local BatchNorm buffers (running_mean 0, running_var 1), never federated."""
    st: Dict[str, torch.Tensor] = {}
    if name == "cifar10_cnn":
        for i, (_, cout) in enumerate(_CIFAR_CONVS, start=1):
            st[f"bn{i}.running_mean"] = torch.zeros(cout)
            st[f"bn{i}.running_var"] = torch.ones(cout)
    return st


# forward ---------------------------------------------------------------------------------


def _drop(x: torch.Tensor, masks: Optional[List[torch.Tensor]], idx: int, p: float, train: bool):
    """Dropout with an INJECTED keep-mask (1/0) so parity does not depend on torch's RNG;
    p == 0 or eval -> identity (nn.Dropout semantics: scale kept values by 1/(1-p))."""
    if not train or p == 0.0 or masks is None:
        return x
    return x * masks[idx].to(x.dtype) / (1.0 - p)


def simple_cnn_forward(w: Dict[str, torch.Tensor], x: torch.Tensor, train: bool = True,
                       dropout_rate: float = 0.0, masks: Optional[List[torch.Tensor]] = None):
    # models_pytorch.py:82-97
    x = F.max_pool2d(F.relu(F.conv2d(x, w["conv1.weight"], w["conv1.bias"], padding=1)), 2, 2)
    x = F.max_pool2d(F.relu(F.conv2d(x, w["conv2.weight"], w["conv2.bias"], padding=1)), 2, 2)
    x = x.reshape(-1, 64 * 7 * 7)
    x = F.relu(F.linear(x, w["fc1.weight"], w["fc1.bias"]))
    x = _drop(x, masks, 0, dropout_rate, train)
    return F.linear(x, w["fc2.weight"], w["fc2.bias"])


def cifar10_cnn_forward(w: Dict[str, torch.Tensor], x: torch.Tensor, train: bool = True,
                        dropout_rate: float = 0.0, masks: Optional[List[torch.Tensor]] = None,
                        bn_state: Optional[Dict[str, torch.Tensor]] = None, eps: float = 1e-5,
                        momentum: float = 0.1, bn_record: Optional[dict] = None, bn_fixed: Optional[dict] = None):
    # models_pytorch.py:136-165; BatchNorm2d defaults eps=1e-5, momentum=0.1, batch stats in train mode.
    # bn_record / bn_fixed (per-sample DP-SGD restatement, oracle/dpsgd.py -- no upstream counterpart): record every
    # layer's batch (mean, biased variance) / normalise with the given constants instead of the batch's own statistics.
    def block(x, i):
        x = F.conv2d(x, w[f"conv{i}.weight"], w[f"conv{i}.bias"], padding=1)
        if bn_fixed is not None:
            m, v = bn_fixed[i]
            return F.relu(F.batch_norm(x, m, v, w[f"bn{i}.weight"], w[f"bn{i}.bias"], False, 0.0, eps))
        if bn_record is not None:
            bn_record[i] = (x.mean((0, 2, 3)).detach(), x.var((0, 2, 3), unbiased=False).detach())
        rm = bn_state[f"bn{i}.running_mean"] if bn_state is not None else None
        rv = bn_state[f"bn{i}.running_var"] if bn_state is not None else None
        use_batch = train or rm is None
        x = F.batch_norm(x, rm, rv, w[f"bn{i}.weight"], w[f"bn{i}.bias"], use_batch, momentum, eps)
        return F.relu(x)

    d = 0
    for a, b in ((1, 2), (3, 4), (5, 6)):
        x = block(block(x, a), b)
        x = F.max_pool2d(x, 2, 2)
        x = _drop(x, masks, d, dropout_rate, train)
        d += 1
    x = x.reshape(-1, 128 * 4 * 4)
    x = _drop(F.relu(F.linear(x, w["fc1.weight"], w["fc1.bias"])), masks, 3, dropout_rate, train)
    x = _drop(F.relu(F.linear(x, w["fc2.weight"], w["fc2.bias"])), masks, 4, dropout_rate, train)
    return F.linear(x, w["fc3.weight"], w["fc3.bias"])


def forward(name: str, w, x, **kw):
    if name == "simple_cnn":
        kw.pop("bn_state", None)
        return simple_cnn_forward(w, x, **kw)
    return cifar10_cnn_forward(w, x, **kw)
