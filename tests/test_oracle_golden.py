"""CPU: the oracle restatement vs golden vectors produced by the unmodified reference
(oracle/make_golden.py).  This is what pins the oracle; the GPU parity tests then compare the
CUDA path with the oracle."""
import numpy as np
import pytest
import torch

from conftest import digest_check, load_golden
from oracle import compression as OC
from oracle import fedavg as OF
from oracle import models as OM
from oracle import philox as OP
from oracle import privacy as OPV
from oracle import round as OR
from oracle import training as OT

torch.set_num_threads(1)


def _batches(model, seed, n, bs):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((n,) + OM.input_shape(model), generator=g)
    y = torch.randint(0, 10, (n,), generator=g)
    return x, y, [(x[i:i + bs], y[i:i + bs]) for i in range(0, n, bs)]


@pytest.mark.parametrize("model", ["simple_cnn", "cifar10_cnn"])
def test_forward_and_grads_match_reference(model):
    gold = load_golden(f"forward_{model}.npz")
    w = OM.init_weights(model, 11)
    x, y, _ = _batches(model, 21, 6, 6)
    assert abs(float(x.double().sum()) - float(gold["x_sum"])) < 1e-9, "seeded input drifted"
    assert np.array_equal(y.numpy(), gold["y"])
    bn = OM.new_bn_state(model)
    loss, logits, grads = OT.loss_and_grads(model, w, x, y, train=True, dropout_rate=0.0, bn_state=bn)
    np.testing.assert_allclose(logits.numpy(), gold["logits"], rtol=1e-5, atol=1e-6)
    assert abs(float(loss) - float(gold["loss"])) < 1e-6
    digest_check(gold, "grad", grads, rtol=1e-4, atol=1e-7)
    ev = OM.forward(model, w, x, train=False, bn_state=bn)
    np.testing.assert_allclose(ev.numpy(), gold["logits_eval"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("model,n,bs", [("simple_cnn", 40, 8), ("cifar10_cnn", 24, 8)])
@pytest.mark.parametrize("opt", ["adam", "sgd", "adamw"])
def test_local_training_matches_reference(model, n, bs, opt):
    gold = load_golden(f"train_{model}_{opt}.npz")
    w = OM.init_weights(model, 12)
    x, y, batches = _batches(model, 22, n, bs)
    assert abs(float(x.double().sum()) - float(gold["x_sum"])) < 1e-9
    bn = OM.new_bn_state(model)
    loss, acc, epochs, samples = OT.train_local_model(model, w, batches, 2, 1e-3 if opt != "sgd" else 1e-2, opt,
                                                      bn_state=bn)
    g_loss, g_acc, g_ep, g_n = gold["metrics"]
    assert (epochs, samples) == (int(g_ep), int(g_n))
    assert abs(loss - g_loss) < 2e-5 and abs(acc - g_acc) < 1e-9
    # same ATen kernels underneath; the only difference is the optimizer written out by hand
    digest_check(gold, "w", w, rtol=2e-4, atol=2e-6)
    if model == "cifar10_cnn":
        digest_check(gold, "buf", bn, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("tag", ["big", "small"])
def test_update_level_dp_matches_reference(tag):
    gold = load_golden("privacy_update_level.npz")
    names = ["a.weight", "a.bias", "b.weight", "b.bias"]
    g = {k: torch.from_numpy(gold[f"{tag}/g/{k}"]) for k in names}
    z = {k: torch.from_numpy(gold[f"{tag}/z/{k}"]) for k in names}
    clipped, norm = OPV.clip(g, 1.0)
    assert norm == float(gold[f"{tag}/norm"])
    for k in names:
        assert np.array_equal(clipped[k].numpy(), gold[f"{tag}/clipped/{k}"]), k     # bit-exact
    noisy, sens, sigma = OPV.add_noise(g, 1.0, 1e-5, 1.0, z)
    assert (sens, sigma) == tuple(gold[f"{tag}/sens_sigma"])
    for k in names:
        assert np.array_equal(noisy[k].numpy(), gold[f"{tag}/noisy/{k}"]), k         # bit-exact
    # reference budget semantics (privacy.py:292,304): one call at (eps, delta) exhausts the budget
    assert tuple(gold[f"{tag}/remaining"]) == (0.0, 0.0)
    assert (tag == "big") == (OPV.global_norm(g) > 1.0)


def test_fedavg_bit_exact_with_reference():
    gold = load_golden("fedavg.npz")
    names = ["l1.weight", "l1.bias", "l2.weight", "l2.bias"]
    K = len(gold["num_samples"])
    ups = [{k: gold[f"theta/{i}/{k}"] for k in names} for i in range(K)]
    ns = [int(v) for v in gold["num_samples"]]
    losses = [float(v) for v in gold["losses"]]
    out = OF.weighted_average(ups, OF.sample_weights(ns))
    for k in names:
        assert np.array_equal(out[k], gold[f"by_samples/{k}"]), k
    out = OF.weighted_average(ups, OF.normalize_weights(list(gold["custom_weights"])))
    for k in names:
        assert np.array_equal(out[k], gold[f"custom/{k}"]), k
    # flat path + max_clients truncation (stable sort by samples, descending)
    theta = np.stack([np.concatenate([u[k].reshape(-1) for k in names]) for u in ups])
    flat, avg_loss, idx, _ = OF.aggregate(theta, ns, losses, min_clients=2, max_clients=4)
    assert idx == [int(v) for v in gold["top4_participants"]]
    ref = np.concatenate([gold[f"top4/{k}"].reshape(-1) for k in names])
    assert np.array_equal(flat, ref)
    assert abs(avg_loss - float(gold["avg_loss"])) < 1e-12


def test_codecs_match_reference():
    gold = load_golden("codecs.npz")
    x = gold["x"]
    for bits, sym in ((8, True), (4, True), (16, True), (8, False)):
        tag = f"q{bits}{'s' if sym else 'a'}"
        q, scale, zp = OC.quantize(x, bits, sym)
        g_scale, g_zp = gold[f"{tag}/scale_zp"]
        assert (scale, zp) == (float(g_scale), int(g_zp))
        assert q.dtype == gold[f"{tag}/q"].dtype
        assert np.array_equal(q, gold[f"{tag}/q"]), tag                                # integer codes bit-exact
        assert np.array_equal(OC.dequantize(q, scale, zp), gold[f"{tag}/dq"]), tag
    for sp in (0.9, 0.5, 0.99995):
        vals, idx = OC.sparsify(x, sp)
        assert len(idx) == len(gold[f"topk{sp}/idx"])
        assert set(idx.tolist()) == set(gold[f"topk{sp}/idx"].tolist())
        assert np.array_equal(OC.desparsify(vals, idx, x.shape), gold[f"topk{sp}/dense"])


def test_round_matches_reference():
    gold = load_golden("round_simple_cnn.npz")
    model = "simple_cnn"
    spec = OM.model_spec(model)
    w0 = OM.init_weights(model, 13)
    gen = torch.Generator().manual_seed(61)
    thetas, ns, losses = [], [], []
    for c in range(3):
        x, y = OR.synthetic_client_data(model, c, n=64 + 32 * c)
        # the generator yields z AFTER training in make_golden; shapes only depend on the spec
        w, loss, acc, n = OR.client_round(model, w0, x, y, dp=False)
        z = {k: torch.randn(spec[k], generator=gen) * 1e-3 for k in spec}
        w, sens, sigma = OPV.apply_update_dp(w, w0, 1.0, 1e-5, 1.0, z)
        g_loss, g_acc, g_n = gold[f"client{c}/metrics"]
        assert n == int(g_n) and abs(loss - g_loss) < 2e-5 and abs(acc - g_acc) < 1e-9
        np.testing.assert_allclose([sens, sigma], gold[f"client{c}/sens_sigma"], rtol=1e-4)
        thetas.append(OR.flatten(w, list(spec)))
        ns.append(n)
        losses.append(loss)
    flat, _, _, _ = OF.aggregate(np.stack(thetas), ns, losses)
    digest_check(gold, "global", OR.unflatten(flat, spec), rtol=2e-4, atol=2e-6)


def test_philox_known_answers():
    for ctr, key, exp in OP.KAT:
        out = OP.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert [int(v) for v in out] == list(exp)


def test_philox_normals_distribution():
    z = OP.normals(2_000_000, seed=42, stream=3).astype(np.float64)
    assert abs(z.mean()) < 3e-3 and abs(z.std() - 1) < 3e-3
    assert abs((z ** 3).mean()) < 1e-2 and abs((z ** 4).mean() - 3) < 3e-2
    # the reference's own statistical window (src/validation/privacy_validator.py:104-108): mean|noise|/sigma in [0.5, 2]
    assert 0.5 <= np.abs(z).mean() <= 2.0
    # streams are independent of how the range is split
    a = OP.raw_blocks(8, 42, 3, first_block=0)
    b = OP.raw_blocks(4, 42, 3, first_block=4)
    assert np.array_equal(a[4:], b)
