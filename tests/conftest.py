import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
STRIDE = 97


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def digest_check(gold, prefix, weights, rtol, atol):
    """Compare a dict of tensors with the strided-sample / sum / l2 digests written by oracle/make_golden.py."""
    import torch
    for k, t in weights.items():
        a = t.detach().reshape(-1).to(torch.float32).cpu().numpy() if hasattr(t, "detach") else np.asarray(t).reshape(-1)
        ref = gold[f"{prefix}/{k}/sample"]
        got = a[::STRIDE] if a.size > 4096 else a
        np.testing.assert_allclose(got, ref, rtol=rtol, atol=atol, err_msg=f"{prefix}/{k}")
        l2 = float(np.sqrt((a.astype(np.float64) ** 2).sum()))
        assert abs(l2 - float(gold[f"{prefix}/{k}/l2"])) <= max(rtol * float(gold[f"{prefix}/{k}/l2"]), atol * np.sqrt(a.size)), k


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
