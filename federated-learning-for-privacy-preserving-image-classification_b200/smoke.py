"""__graft_entry__.smoke(): one small federated round on cuda:0 checked against the CPU oracle.
(The only product-tree file allowed to import ``oracle`` -- it is the checker here, not the path.)"""
from __future__ import annotations

import numpy as np
import torch


def run(device: torch.device) -> None:
    from oracle import models as OM
    from oracle import round as OR
    from .simulation import FederatedRoundEngine

    model, K, sizes = "simple_cnn", 3, [40, 33, 64]
    w0 = OM.init_weights(model, 13)
    data = [OR.synthetic_client_data(model, c, n=sizes[c]) for c in range(K)]
    gen = torch.Generator().manual_seed(61)
    spec = OM.model_spec(model)
    zs = [{k: torch.randn(spec[k], generator=gen) * 1e-3 for k in spec} for _ in range(K)]
    ref, info = OR.federated_round(model, w0, K, dp=True, zs=zs, data=data, batch_size=8, lr=1e-2, optimizer="sgd")

    eng = FederatedRoundEngine(model, K, device, batch_size=8, learning_rate=1e-2, optimizer_type="sgd",
                               dp_mode="update", dropout_rate=0.0, precision="fp32")
    eng.set_global_weights(w0)
    eng.load_data([d[0] for d in data], [d[1] for d in data], sizes)
    zrows = eng.layout.new_rows(K, eng.device)
    for k in range(K):
        eng.layout.flatten_into(zrows[k], zs[k])
    eng.dp_z = zrows
    out = eng.run_round()
    got = eng.global_weights("cpu")
    for name in ref:
        np.testing.assert_allclose(got[name].numpy(), ref[name].numpy(), rtol=2e-4, atol=2e-6, err_msg=name)
    assert out["samples"] == sizes

    # the same round on the tcgen05 TF32 path (TMA-fed tensor-core convolutions / linears) at its stated tolerance:
    # the accumulated update within 5 % relative L2 of the oracle's
    eng2 = FederatedRoundEngine(model, K, device, batch_size=8, learning_rate=1e-2, optimizer_type="sgd",
                                dp_mode="update", dropout_rate=0.0, precision="tf32")
    eng2.set_global_weights(w0)
    eng2.load_data([d[0] for d in data], [d[1] for d in data], sizes)
    eng2.dp_z = zrows
    eng2.run_round()
    got2 = eng2.global_weights("cpu")
    num = sum(float(((got2[n] - ref[n]) ** 2).sum()) for n in ref)
    den = sum(float(((ref[n] - w0[n]) ** 2).sum()) for n in ref)
    assert (num / den) ** 0.5 < 5e-2, (num / den) ** 0.5
    print(f"smoke ok: round of {K} clients matches the oracle (fp32 path exact-order tolerance, TF32 path "
          f"{(num / den) ** 0.5:.1e} rel. L2 of the update); losses {['%.4f' % l for l in out['losses']]}")
