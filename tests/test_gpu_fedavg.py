"""GPU parity: csrc/fedavg.cu through the C ABI vs the oracle (oracle/fedavg.py) and the golden vectors."""
from datetime import datetime

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import compression as OC
from oracle import fedavg as OF

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,P", [(1, 1), (2, 31), (3, 1000), (10, 421642), (50, 421642), (7, 4097), (1100, 2051)])
def test_flat_bit_exact_vs_oracle(cuda_device, K, P):
    from flb200 import ops
    from flb200.layout import ParamLayout
    rng = np.random.default_rng(K * 1000 + P)
    theta = (rng.standard_normal((K, P)) * 0.01).astype(np.float32)
    ns = rng.integers(100, 1000, K).tolist()
    w = OF.sample_weights(ns)
    ld = (P + 31) // 32 * 32
    dev = torch.zeros((K, ld), dtype=torch.float32, device=cuda_device)
    dev[:, :P] = torch.from_numpy(theta).to(cuda_device)
    out = ops.fedavg_weighted_sum(dev, w, P=P)
    ref = OF.weighted_average_flat(theta, w)
    assert np.array_equal(out.cpu().numpy(), ref)                       # bit-exact (fp32 mul, fp32 add, client order)
    # unaligned rows take the scalar kernel; same bits
    un = torch.from_numpy(theta).to(cuda_device).contiguous()
    assert np.array_equal(ops.fedavg_weighted_sum(un, w).cpu().numpy(), ref)
    # accumulate=True continues a chunked aggregation: first half, then second half
    if K >= 2:
        h = K // 2
        part = ops.fedavg_weighted_sum(dev[:h], w[:h], P=P)
        ops.fedavg_weighted_sum(dev[h:], w[h:], P=P, out=part, accumulate=True)
        assert np.array_equal(part.cpu().numpy(), ref)


def test_empty_and_errors(cuda_device):
    from flb200 import ops
    from flb200._lib import FlbError
    z = ops.fedavg_weighted_sum(torch.zeros((3, 0), device=cuda_device), [0.2, 0.3, 0.5])
    assert z.numel() == 0
    with pytest.raises(FlbError):
        ops.fedavg_weighted_sum(torch.zeros((3, 32), device=cuda_device), [0.5, 0.5])
    with pytest.raises(FlbError):
        ops.fedavg_weighted_sum(torch.zeros((3, 32)), [0.2, 0.3, 0.5])          # host tensor: no CPU path


def _updates(gold, names, device):
    from flb200.models import ModelUpdate
    K = len(gold["num_samples"])
    return [ModelUpdate(client_id=f"c{i}", round_number=3,
                        model_weights={k: torch.from_numpy(gold[f"theta/{i}/{k}"]).to(device) for k in names},
                        num_samples=int(gold["num_samples"][i]), training_loss=float(gold["losses"][i]),
                        privacy_budget_used=0.5, compression_ratio=0.8, timestamp=datetime.now()) for i in range(K)]


@pytest.mark.parametrize("where", ["cpu", "cuda"])
def test_aggregator_matches_reference_golden(cuda_device, where):
    """The drop-in class against outputs of the unmodified reference FedAvgAggregator (tests/golden/fedavg.npz)."""
    from flb200.fedavg import FedAvgAggregator, FedAvgError
    gold = load_golden("fedavg.npz")
    names = ["l1.weight", "l1.bias", "l2.weight", "l2.bias"]
    dev = torch.device("cpu") if where == "cpu" else cuda_device
    ups = _updates(gold, names, dev)
    agg = FedAvgAggregator(min_clients=2, validate_updates=False)
    gm = agg.aggregate_updates(ups)
    assert gm.round_number == 3 and gm.participating_clients == [u.client_id for u in ups]
    for k in names:
        assert gm.model_weights[k].device.type == where
        assert np.array_equal(gm.model_weights[k].cpu().numpy(), gold[f"by_samples/{k}"]), k
    gm = agg.aggregate_updates(ups, weights=[float(v) for v in gold["custom_weights"]])
    for k in names:
        assert np.array_equal(gm.model_weights[k].cpu().numpy(), gold[f"custom/{k}"]), k
    gm = FedAvgAggregator(min_clients=2, max_clients=4, validate_updates=False).aggregate_updates(ups)
    assert gm.participating_clients == [f"c{int(i)}" for i in gold["top4_participants"]]
    for k in names:
        assert np.array_equal(gm.model_weights[k].cpu().numpy(), gold[f"top4/{k}"]), k
    assert abs(agg.aggregation_history[0]["avg_training_loss"] - float(
        sum(float(l) * w for l, w in zip(gold["losses"], OF.sample_weights([int(v) for v in gold["num_samples"]]))))) < 1e-12
    # error behaviour of the reference: wrapped message, insufficient updates, bad weights
    with pytest.raises(FedAvgError, match="FedAvg aggregation failed: Insufficient valid updates: 1 < 2"):
        agg.aggregate_updates(ups[:1])
    with pytest.raises(FedAvgError, match="No model updates provided"):
        agg.aggregate_updates([])
    with pytest.raises(FedAvgError, match="Number of weights must match"):
        agg.aggregate_updates(ups, weights=[1.0])
    # validator on: weights above the 10.0 magnitude limit are dropped (validation.py:24,87-91)
    big = _updates(gold, names, dev)
    big[0].model_weights["l1.bias"] = big[0].model_weights["l1.bias"] + 100.0
    gm = FedAvgAggregator(min_clients=2, validate_updates=True).aggregate_updates(big)
    assert "c0" not in gm.participating_clients


def test_q8_fused_dequant_average(cuda_device):
    from flb200 import ops
    from flb200.layout import ParamLayout
    from collections import OrderedDict
    lay = ParamLayout(OrderedDict([("a.weight", (37, 19)), ("a.bias", (37,)), ("b.weight", (5, 37)), ("b.bias", (5,))]))
    K = 6
    rng = np.random.default_rng(5)
    x = rng.standard_normal((K, lay.P)).astype(np.float32)
    w = OF.sample_weights(rng.integers(100, 1000, K).tolist())
    rows = lay.new_rows(K, cuda_device)
    rows[:, :lay.P] = torch.from_numpy(x).to(cuda_device)
    seg = lay.seg_off(cuda_device)
    q, scale, zp = ops.q8_quantize(rows, seg, P=lay.P)
    # oracle: per (client, layer) quantise -> dequantise -> sequential fp32 FedAvg
    deq = np.zeros_like(x)
    offs = seg.cpu().numpy()
    for k in range(K):
        for l in range(len(lay.names)):
            sl = slice(offs[l], offs[l + 1])
            qq, s, z = OC.quantize(x[k, sl])
            assert np.array_equal(q[k, sl].cpu().numpy(), qq), (k, l)               # integer codes bit-exact
            assert np.float32(s) == scale[k, l].item() and z == int(zp[k, l].item())
            deq[k, sl] = OC.dequantize(qq, s, z)
    got = ops.q8_dequantize(q, scale, zp, seg, P=lay.P)[:, :lay.P].cpu().numpy()
    assert np.array_equal(got, deq)
    out = ops.fedavg_weighted_sum_q8(q, scale, zp, seg, w, lay.P)
    assert np.array_equal(out.cpu().numpy(), OF.weighted_average_flat(deq, w))


def test_full_size_linearity_property(cuda_device):
    """BASELINE config 5 corner (K=100, P=10M): FedAvg is linear, so avg(a*theta) == a*avg(theta) for a power of two,
    and a permutation of clients with permuted weights only changes the rounding order (<= 1e-6 * max)."""
    from flb200 import ops
    K, P = 100, 10_000_000
    g = torch.Generator(device=cuda_device).manual_seed(7)
    theta = torch.randn((K, P), generator=g, device=cuda_device) * 0.01
    n = torch.randint(100, 1000, (K,), generator=torch.Generator().manual_seed(7)).tolist()
    w = OF.sample_weights(n)
    base = ops.fedavg_weighted_sum(theta, w)
    assert torch.equal(ops.fedavg_weighted_sum(theta * 4.0, w), base * 4.0)
    perm = torch.randperm(K, generator=torch.Generator().manual_seed(1)).tolist()
    other = ops.fedavg_weighted_sum(theta[perm].contiguous(), [w[i] for i in perm])
    assert (other - base).abs().max().item() <= 1e-6 * theta.abs().max().item()
    # spot-check 1000 random columns against the oracle
    cols = torch.randint(0, P, (1000,), generator=torch.Generator().manual_seed(2))
    ref = OF.weighted_average_flat(theta[:, cols.to(cuda_device)].cpu().numpy(), w)
    assert np.array_equal(base[cols.to(cuda_device)].cpu().numpy(), ref)


@pytest.mark.gpu
def test_benchmark_helper_report_shape(cuda_device):
    """fedavg.py:487-546: same keys per client count, memory = 4 layers x (size // 4) fp32."""
    from flb200.fedavg import benchmark_aggregation_performance
    r = benchmark_aggregation_performance([3, 6], model_size=40000, device=cuda_device)
    assert sorted(r) == ["3_clients", "6_clients"]
    for n in (3, 6):
        row = r[f"{n}_clients"]
        assert row["participating_clients"] == n and row["memory_usage"] == 40000 * 4
        assert row["aggregation_time"] > 0 and abs(row["throughput"] * row["aggregation_time"] - n) < 1e-9
    assert "error" in benchmark_aggregation_performance([1], model_size=400, device=cuda_device)["1_clients"]     # < min_clients
