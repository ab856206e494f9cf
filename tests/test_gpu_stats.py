"""GPU parity of the fused update validation / convergence reductions against the reference's
per-tensor formulas (src/shared/validation.py:72-91, src/aggregation/fedavg.py:144-190)."""
from datetime import datetime

import numpy as np
import pytest
import torch

from oracle import models as OM

pytestmark = pytest.mark.gpu


def _update(w, **kw):
    from flb200.models import ModelUpdate
    return ModelUpdate(kw.get("cid", "client_0"), 1, w, 100, 0.5, 0.1, 0.8, datetime.now())


def test_update_stats_kernel_batched(cuda_device):
    from flb200 import ops
    from flb200.layout import ParamLayout
    lay = ParamLayout(OM.model_spec("simple_cnn"))
    K = 5
    rows = lay.new_rows(K, cuda_device)
    g = torch.Generator().manual_seed(3)
    rows[:, :lay.P] = (torch.randn((K, lay.P), generator=g) * 0.1).to(cuda_device)
    rows[1, lay.offsets["fc1.weight"] + 12345] = float("nan")
    rows[2, lay.offsets["conv2.bias"] + 3] = float("-inf")
    rows[3, lay.offsets["fc2.weight"] + 7] = -42.5
    offs = [lay.offsets[n] for n in lay.names]
    mx, fl = ops.update_stats(ops.rows_ptr_table(rows, offs), lay.seg_off(cuda_device), K, lay.P, cuda_device)
    mx, fl = mx.cpu().numpy(), fl.cpu().numpy()
    for k in range(K):
        for l, name in enumerate(lay.names):
            t = rows[k, lay.offsets[name]:lay.offsets[name] + int(np.prod(lay.shapes[name]))].cpu()
            assert bool(fl[k, l] & 1) == bool(torch.isnan(t).any()), (k, name)
            assert bool(fl[k, l] & 2) == bool(torch.isinf(t).any()), (k, name)
            ref = float(torch.where(torch.isnan(t), torch.zeros_like(t), t.abs()).max())
            assert mx[k, l] == np.float32(ref), (k, name)


def test_validator_messages_match_host_path(cuda_device):
    from flb200.validation import ModelUpdateValidator, ValidationError
    v = ModelUpdateValidator()
    w = OM.init_weights("cifar10_cnn", 1)
    good = {k: t.to(cuda_device) for k, t in w.items()}
    assert v.validate_model_update(_update(good)) and v.validate_model_update(_update(w))
    for poke, msg in ((float("nan"), "NaN values found in layer conv3.weight"),
                      (float("inf"), "Infinite values found in layer conv3.weight"),
                      (11.0, "exceeds maximum 10.0 in layer conv3.weight")):
        bad = {k: t.clone() for k, t in w.items()}
        bad["conv3.weight"].view(-1)[17] = poke
        errs = []
        for dev_w in (bad, {k: t.to(cuda_device) for k, t in bad.items()}):        # host path, fused device path
            with pytest.raises(ValidationError) as e:
                v.validate_model_update(_update(dev_w))
            errs.append(str(e.value))
        assert msg in errs[0] and errs[0] == errs[1]


def test_convergence_metric_fused_matches_formula(cuda_device):
    from flb200.fedavg import FedAvgAggregator
    from flb200.models import GlobalModel
    old = OM.init_weights("simple_cnn", 1)
    new = {k: t + 0.01 * torch.randn(t.shape, generator=torch.Generator().manual_seed(2)) for k, t in old.items()}
    ref_d = sum(float(torch.norm(new[k] - old[k])) for k in new)
    ref_n = sum(float(torch.norm(new[k])) for k in new)
    agg = FedAvgAggregator(validate_updates=False)
    mk = lambda w: GlobalModel(1, w, {}, ["a"], 0.0, datetime.now())
    got = agg.calculate_convergence_metrics(mk({k: t.to(cuda_device) for k, t in old.items()}),
                                            mk({k: t.to(cuda_device) for k, t in new.items()}))
    assert abs(got - ref_d / ref_n) < 1e-6
    assert abs(agg.calculate_convergence_metrics(mk(old), mk(new)) - ref_d / ref_n) < 1e-6      # host tensors: torch path
