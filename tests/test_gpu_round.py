"""GPU parity: one full federated round (train -> update-level DP -> FedAvg) vs the oracle and the reference's golden."""
import numpy as np
import pytest
import torch

from conftest import digest_check, load_golden
from oracle import fedavg as OF
from oracle import models as OM
from oracle import privacy as OPV
from oracle import round as OR

pytestmark = pytest.mark.gpu
MODEL = "simple_cnn"


def _zrows(eng, zs):
    rows = eng.layout.new_rows(len(zs), eng.device)
    for k, z in enumerate(zs):
        eng.layout.flatten_into(rows[k], z)
    return rows


def test_round_matches_reference_golden(cuda_device):
    """tests/golden/round_simple_cnn.npz was produced by the unmodified reference classes (LocalTrainer, Adam 1e-3,
    batch 32 -> DifferentialPrivacyEngine with injected noise -> FedAvgAggregator)."""
    from flb200.simulation import FederatedRoundEngine
    gold = load_golden("round_simple_cnn.npz")
    spec = OM.model_spec(MODEL)
    w0 = OM.init_weights(MODEL, 13)
    sizes = [64, 96, 128]
    data = [OR.synthetic_client_data(MODEL, c, n=sizes[c]) for c in range(3)]
    gen = torch.Generator().manual_seed(61)
    zs = [{k: torch.randn(spec[k], generator=gen) * 1e-3 for k in spec} for _ in range(3)]
    eng = FederatedRoundEngine(MODEL, 3, cuda_device, dp_mode="update", dropout_rate=0.0, precision="fp32")
    eng.set_global_weights(w0)
    eng.load_data([d[0] for d in data], [d[1] for d in data], sizes)
    eng.dp_z = _zrows(eng, zs)
    out = eng.run_round()
    for c in range(3):
        g_loss, g_acc, g_n = gold[f"client{c}/metrics"]
        assert out["samples"][c] == int(g_n) and abs(out["losses"][c] - g_loss) < 1e-4 and abs(out["accuracies"][c] - g_acc) < 1e-9
        sens, sigma = gold[f"client{c}/sens_sigma"]
        assert abs(min(eng.norms[c].item(), 1.0) - sens) < 2e-3 * sens     # sensitivity = min(||delta||, C)
    # global model: Adam steps of +-lr on near-zero-gradient coordinates differ (see conftest.adam_trajectory_check);
    # the aggregated update is compared on its own scale
    got = eng.global_weights("cpu")
    num = den = 0.0
    for k in spec:
        a = got[k].reshape(-1).numpy()
        ref = gold[f"global/{k}/sample"]
        s = a[::97] if a.size > 4096 else a
        b0 = w0[k].reshape(-1).numpy()
        b0 = b0[::97] if b0.size > 4096 else b0
        assert np.abs(s - ref).max() < 2e-3
        num += float(((s - ref) ** 2).sum()); den += float(((ref - b0) ** 2).sum())
    assert (num / den) ** 0.5 < 2e-2


@pytest.mark.parametrize("compression", [None, "q8", "topk"])
def test_round_sgd_vs_oracle_tight(cuda_device, compression):
    """SGD(momentum) has no sign sensitivity: the whole round matches the oracle at fp32 tolerance
    (q8: after the oracle's quantise -> dequantise of each upload)."""
    from flb200.simulation import FederatedRoundEngine
    from oracle import compression as OC
    K, sizes = 4, [40, 33, 64, 7]
    spec = OM.model_spec(MODEL)
    w0 = OM.init_weights(MODEL, 3)
    data = [OR.synthetic_client_data(MODEL, c, n=sizes[c]) for c in range(K)]
    gen = torch.Generator().manual_seed(5)
    zs = [{k: torch.randn(spec[k], generator=gen) * 1e-3 for k in spec} for _ in range(K)]
    ref, info = OR.federated_round(MODEL, w0, K, dp=True, zs=zs, data=data, batch_size=8, lr=1e-2, optimizer="sgd")
    eng = FederatedRoundEngine(MODEL, K, cuda_device, batch_size=8, learning_rate=1e-2, optimizer_type="sgd",
                               dp_mode="update", dropout_rate=0.0, precision="fp32", compression=compression)
    eng.set_global_weights(w0)
    eng.load_data([d[0] for d in data], [d[1] for d in data], sizes)
    eng.dp_z = _zrows(eng, zs)
    out = eng.run_round()
    assert out["samples"] == sizes
    np.testing.assert_allclose(out["losses"], info["losses"], atol=2e-5)
    got = eng.global_weights("cpu")
    if compression == "q8":
        names = list(spec)
        deq = []
        for theta in info["client_thetas"]:
            parts, off = [], 0
            for n in names:
                k = int(np.prod(spec[n]))
                q, s, z = OC.quantize(theta[off:off + k])
                parts.append(OC.dequantize(q, s, z)); off += k
            deq.append(np.concatenate(parts))
        flat = OF.weighted_average_flat(np.stack(deq), info["weights"])
        ref = OR.unflatten(flat, spec)
        for name in ref:      # one code step (scale) where a value sits on a rounding boundary, else exact
            scale = 2 * float(np.abs(ref[name].numpy()).max()) / 255 + 1e-12
            assert np.abs(got[name].numpy() - ref[name].numpy()).max() <= 1.01 * scale, name
            assert np.mean(np.abs(got[name].numpy() - ref[name].numpy()) > 1e-6) < 0.02, name
    elif compression == "topk":
        # TopKSparsificationCompressor (sparsity 0.9) on every upload, then FedAvg of the reconstructed dense rows.  The kept
        # SET is decided on each side's own uploads (fp32 reordering moves entries sitting exactly at the k-th magnitude), so
        # the aggregate is compared entry by entry with a small budget for such boundary swaps.
        names = list(spec)
        dense = []
        for theta in info["client_thetas"]:
            parts, off = [], 0
            for n in names:
                k = int(np.prod(spec[n]))
                v, i = OC.sparsify(theta[off:off + k], 0.9)
                parts.append(OC.desparsify(v, i, (k,))); off += k
            dense.append(np.concatenate(parts))
        flat = OF.weighted_average_flat(np.stack(dense), info["weights"])
        ref = OR.unflatten(flat, spec)
        for name in ref:
            g_, r_ = got[name].numpy().reshape(-1), ref[name].numpy().reshape(-1)
            close = np.abs(g_ - r_) <= 2e-4 * np.abs(r_) + 2e-6
            assert close.mean() > 0.995, (name, close.mean())
            assert (g_ != 0).sum() <= int(r_.size * 0.1 * K) + K      # at most K clients x 10 % of the entries survive
    else:
        for name in ref:
            np.testing.assert_allclose(got[name].numpy(), ref[name].numpy(), rtol=2e-4, atol=2e-6, err_msg=name)
    # second round continues from the aggregated model with fresh optimizer state and a fresh noise stream
    eng.dp_z = None
    out2 = eng.run_round()
    assert out2["round"] == 2 and all(np.isfinite(out2["losses"]))


def test_simulation_signature_and_result_dict(cuda_device):
    from flb200.simulation import FederatedLearningSimulation, SimulationConfig
    cfg = SimulationConfig(num_clients=5, num_rounds=2)
    assert cfg.to_dict()["privacy_epsilon"] == 1.0 and cfg.model_type == "simple_cnn"
    data = [OR.synthetic_client_data(MODEL, c, n=64) for c in range(5)]
    res = FederatedLearningSimulation(cfg, device=cuda_device, local_epochs=1, dp_mode="none",
                                      data=([d[0] for d in data], [d[1] for d in data])).run_simulation(timeout_minutes=5)
    for key in ("simulation_config", "start_time", "end_time", "duration_seconds", "success", "clients", "summary"):
        assert key in res, res
    assert res["success"] and res["summary"]["total_rounds"] == 2 and res["summary"]["total_clients"] == 5
    assert len(res["clients"]["client_0"]["training_history"]) == 2
    bad = FederatedLearningSimulation(SimulationConfig(model_type="nope"), device=cuda_device).run_simulation()
    assert "error" in bad                      # never raises (federated_simulation.py:399-405)


def test_smoke_entry(cuda_device):
    import __graft_entry__ as g
    g.smoke()


def test_prefetched_uploads_match_plain_loads(cuda_device):
    """Double-buffered host -> device path (prefetch_packed / use_prefetched, two captured graphs) gives exactly the
    rounds of the plain load_packed path, with different data every round."""
    from flb200.simulation import FederatedRoundEngine
    K, sizes = 3, [40, 33, 64]
    w0 = OM.init_weights(MODEL, 3)
    rounds = []
    for r in range(4):
        data = [OR.synthetic_client_data(MODEL, 10 * r + c, n=sizes[c]) for c in range(K)]
        x = torch.cat([d[0].reshape(d[0].shape[0], -1) for d in data]).pin_memory()
        y = torch.cat([d[1] for d in data]).to(torch.int32).pin_memory()
        rounds.append((x, y))
    outs = []
    for mode in ("plain", "prefetch"):
        eng = FederatedRoundEngine(MODEL, K, cuda_device, batch_size=8, learning_rate=1e-2, optimizer_type="sgd",
                                   dp_mode="none", dropout_rate=0.0, precision="fp32")
        eng.set_global_weights(w0)
        eng.load_packed(rounds[0][0], rounds[0][1], sizes)
        model_host = torch.zeros(eng.layout.P, dtype=torch.float32).pin_memory()
        losses = []
        for r in range(4):
            if mode == "plain":
                eng.load_packed(rounds[r][0], rounds[r][1], sizes)
            else:
                if r == 0:
                    eng.prefetch_packed(rounds[0][0], rounds[0][1])
                eng.use_prefetched()
                eng.start_round()                      # the split form: the next upload is issued while the round runs
                if r + 1 < 4:
                    eng.prefetch_packed(rounds[r + 1][0], rounds[r + 1][1])
                out = eng.finish_round(model_out=model_host)
                assert torch.equal(model_host[:eng.layout.P], eng.global_row[:eng.layout.P].cpu())
                losses.append(out["losses"])
                continue
            losses.append(eng.run_round()["losses"])
        outs.append((losses, eng.global_row.clone()))
    np.testing.assert_allclose(outs[0][0], outs[1][0], rtol=1e-5)          # split-K atomics reorder fp32 sums between runs
    # atomics reorder fp32 sums between runs: same tolerance as two plain runs
    torch.testing.assert_close(outs[0][1], outs[1][1], rtol=1e-4, atol=1e-6)


def test_evaluate_global_matches_oracle_forward(cuda_device):
    """Held-out evaluation of the global model (training.py:307-360) through the batched forward kernels."""
    from flb200.simulation import FederatedRoundEngine
    w0 = OM.init_weights(MODEL, 9)
    eng = FederatedRoundEngine(MODEL, 2, cuda_device, dp_mode="none", dropout_rate=0.25, precision="fp32")
    eng.set_global_weights(w0)
    x, y = OR.synthetic_client_data(MODEL, 77, n=203)
    got = eng.evaluate_global(x, y, shards=3)
    logits = OM.forward(MODEL, w0, x, train=False)
    assert got["samples"] == 203
    assert abs(got["accuracy"] - float((logits.argmax(1) == y).double().mean())) < 1e-12
