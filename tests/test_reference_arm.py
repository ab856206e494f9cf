"""CPU: the staged reference modules (oracle/_ref, byte-identical copies made by oracle/build_ref.py) drive a full round
through the reference's OWN LocalTrainer / DifferentialPrivacyEngine / FedAvgAggregator, and the oracle port reproduces
that round bit for bit -- which is what makes the port a valid checker and the staged copy a valid CPU baseline."""
import hashlib
import json
import os

import pytest
import torch

from oracle import build_ref
from oracle import models as OM
from oracle import round as OR

pytestmark = pytest.mark.skipif(not build_ref.available(), reason="oracle/_ref not staged (python -m oracle.build_ref)")


def test_manifest_matches_staged_files():
    man = json.load(open(os.path.join(build_ref.DEST, "MANIFEST.json")))
    assert sorted(man["files"]) == sorted(build_ref.MODULES)
    for rel, sha in man["files"].items():
        assert hashlib.sha256(open(os.path.join(build_ref.DEST, rel), "rb").read()).hexdigest() == sha
        ref = os.path.join("/root/reference", rel)
        if os.path.exists(ref):                                   # build container: the copy is the reference, unmodified
            assert open(ref, "rb").read() == open(os.path.join(build_ref.DEST, rel), "rb").read()


def test_reference_round_equals_oracle_port_bit_for_bit():
    from oracle import ref_round as RR
    torch.set_num_threads(1)
    model = "simple_cnn"
    w0 = OM.init_weights(model, 5)
    data = [OR.synthetic_client_data(model, c, n=n) for c, n in enumerate([40, 72])]
    got, info = RR.federated_round(model, w0, data, dp=False, dropout_rate=0.0, batch_size=16)
    ref, rinfo = OR.federated_round(model, w0, 2, dp=False, data=data, dropout_rate=0.0, batch_size=16)
    assert info["num_samples"] == rinfo["num_samples"] == [40, 72]
    for k in ref:
        assert torch.equal(got[k], ref[k]), k
    # with DP on, the reference draws torch.normal noise of std sigma = S * 4.8448: the round still returns finite weights
    # of the right shapes and consumes the budget of a fresh engine per client
    noisy, _ = RR.federated_round(model, w0, data, dp=True, batch_size=16)
    assert all(torch.isfinite(noisy[k]).all() and noisy[k].shape == w0[k].shape for k in w0)
