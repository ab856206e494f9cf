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


def adam_trajectory_check(gold, prefix, weights, w0, lr, rel_l2=5e-3):
    """Adam's update is ~lr*g/(|g|+eps): coordinates whose gradient is within rounding of zero take a different
    +-lr step under ANY change of summation order, so trajectories are compared (SURVEY.md section 7 'hard parts') by
    (a) |w - w_ref| <= 2*lr elementwise and (b) relative L2 error of the accumulated update, on the golden samples."""
    import torch
    num = den = 0.0
    for k, t in weights.items():
        a = t.detach().reshape(-1).to(torch.float32).cpu().numpy()
        b0 = w0[k].detach().reshape(-1).cpu().numpy()
        ref = gold[f"{prefix}/{k}/sample"]
        got, start = (a[::STRIDE], b0[::STRIDE]) if a.size > 4096 else (a, b0)
        assert np.abs(got - ref).max() <= 2 * lr, k
        num += float(((got - ref).astype(np.float64) ** 2).sum())
        den += float(((ref - start).astype(np.float64) ** 2).sum())
    assert (num / den) ** 0.5 <= rel_l2, (num / den) ** 0.5
