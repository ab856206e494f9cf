"""GPU parity: the batched SimpleCNN training kernels (through the C ABI) vs the oracle (oracle/training.py,
oracle/dpsgd.py) and vs golden vectors produced by the unmodified reference LocalTrainer."""
import numpy as np
import pytest
import torch

from conftest import adam_trajectory_check, digest_check, load_golden
from oracle import dpsgd as OD
from oracle import models as OM
from oracle import privacy as OPV
from oracle import training as OT

pytestmark = pytest.mark.gpu
MODEL = "simple_cnn"
# fp32 CUDA-core path: same fp32 arithmetic as the reference, different summation order only
RTOL, ATOL = 2e-4, 2e-6


def _data(seed, n):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((n,) + OM.input_shape(MODEL), generator=g)
    y = torch.randint(0, 10, (n,), generator=g)
    return x, y


def _engine(cuda_device, K, B, **kw):
    from flb200.training import BatchedClientTrainer
    kw.setdefault("dropout_rate", 0.0)
    kw.setdefault("precision", "fp32")
    return BatchedClientTrainer(MODEL, K, cuda_device, batch_size=B, **kw)


def test_forward_and_grads_match_reference_golden(cuda_device):
    gold = load_golden("forward_simple_cnn.npz")
    w = OM.init_weights(MODEL, 11)
    x, y = _data(21, 6)
    eng = _engine(cuda_device, 1, 6)
    eng.set_client_weights(0, w)
    eng.load_data([x], [y])
    eng.forward_backward()
    logits = eng.ws_array("logits", torch.float32, 10)[0, :6].cpu().numpy()
    np.testing.assert_allclose(logits, gold["logits"], rtol=1e-4, atol=1e-5)
    loss, acc, seen = eng.epoch_metrics()
    assert abs(float(loss[0]) - float(gold["loss"])) < 1e-5 and int(seen[0]) == 6
    digest_check(gold, "grad", eng.layout.views(eng.G[0]), rtol=1e-3, atol=1e-6)
    # eval-mode forward (model.forward / evaluate_model path)
    _, _, _, lg = eng.evaluate()
    np.testing.assert_allclose(lg.cpu().numpy(), gold["logits_eval"], rtol=1e-4, atol=1e-5)


def test_batched_ragged_clients_grads_vs_oracle(cuda_device):
    """3 clients with different weights and different (ragged) batch sizes in one launch sequence."""
    sizes = [32, 17, 1]
    eng = _engine(cuda_device, 3, 32)
    ws, xs, ys = [], [], []
    for k, n in enumerate(sizes):
        ws.append(OM.init_weights(MODEL, 30 + k))
        x, y = _data(40 + k, n)
        xs.append(x); ys.append(y)
        eng.set_client_weights(k, ws[k])
    eng.load_data(xs, ys)
    eng.forward_backward()
    for k, n in enumerate(sizes):
        loss, logits, grads = OT.loss_and_grads(MODEL, ws[k], xs[k], ys[k], train=True, dropout_rate=0.0)
        got = eng.layout.views(eng.G[k])
        for name, g in grads.items():
            ref = g.numpy()
            np.testing.assert_allclose(got[name].cpu().numpy(), ref, rtol=2e-3, atol=2e-5 * max(1e-3, np.abs(ref).max()), err_msg=f"{k}/{name}")
        np.testing.assert_allclose(eng.ws_array("logits", torch.float32, 10)[k, :n].cpu().numpy(), logits.numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("opt", ["adam", "sgd", "adamw"])
def test_training_trajectory_matches_reference_golden(cuda_device, opt):
    """2 epochs x 5 steps of the unmodified reference LocalTrainer (golden) vs the kernels."""
    gold = load_golden(f"train_simple_cnn_{opt}.npz")
    w = OM.init_weights(MODEL, 12)
    x, y = _data(22, 40)
    eng = _engine(cuda_device, 1, 8)
    eng.set_client_weights(0, w)
    eng.load_data([x], [y])
    loss, acc, samples = eng.train(2, 1e-3 if opt != "sgd" else 1e-2, opt)
    g_loss, g_acc, g_ep, g_n = gold["metrics"]
    assert samples[0] == int(g_n)
    assert abs(loss[0] - g_loss) < 1e-4 and abs(acc[0] - g_acc) < 1e-9
    if opt == "sgd":
        digest_check(gold, "w", eng.client_weights(0, "cpu"), rtol=RTOL, atol=ATOL)
    else:
        adam_trajectory_check(gold, "w", eng.client_weights(0, "cpu"), w, 1e-3)


def test_multi_client_epoch_graph_matches_oracle(cuda_device):
    """4 clients, ragged sample counts (different step counts per client), 3 epochs: eager, captured and replayed
    epochs all follow the oracle; a client that runs out of batches stops stepping."""
    sizes = [40, 33, 64, 7]
    w0 = OM.init_weights(MODEL, 5)
    xs, ys = zip(*[_data(70 + k, n) for k, n in enumerate(sizes)])
    eng = _engine(cuda_device, 4, 8)
    eng.set_global_row(eng.layout.flatten(w0, cuda_device))
    eng.load_data(xs, ys)
    loss, acc, samples = eng.train(3, 1e-3, "adam")
    assert eng._graph is not None                       # third epoch was a graph replay
    assert samples == [3 * n for n in sizes]
    for k, n in enumerate(sizes):
        w = {a: b.clone() for a, b in w0.items()}
        l, a, ep, tot = OT.train_local_model(MODEL, w, OT.make_batches(xs[k], ys[k], 8), 3, 1e-3, "adam")
        assert abs(loss[k] - l) < 1e-4 and abs(acc[k] - a) < 1e-9, k
        got = eng.client_weights(k, "cpu")
        for name in w:
            # Adam's first steps are ~lr*sign(g): compare on the scale of the accumulated update
            np.testing.assert_allclose(got[name].numpy(), w[name].numpy(), rtol=0, atol=2e-3, err_msg=f"{k}/{name}")
        upd = torch.cat([(got[n_] - w0[n_]).flatten() for n_ in w])
        ref = torch.cat([(w[n_] - w0[n_]).flatten() for n_ in w])
        assert float((upd - ref).norm() / ref.norm()) < 5e-3


def test_local_trainer_drop_in(cuda_device):
    """Reference call signature: LocalTrainer(model, device).train_local_model(loader, epochs, lr, 'adam', ...)."""
    from torch.utils.data import DataLoader, TensorDataset
    from flb200.models_pytorch import ModelFactory
    from flb200.training import LocalTrainer, TrainingError
    gold = load_golden("train_simple_cnn_adam.npz")
    model = ModelFactory.create_model("simple_cnn", dropout_rate=0.0)
    assert model.get_parameter_count() == 421642
    model.set_model_weights(OM.init_weights(MODEL, 12))
    x, y = _data(22, 40)
    loader = DataLoader(TensorDataset(x, y), batch_size=8, shuffle=False)
    trainer = LocalTrainer(model, cuda_device)
    m = trainer.train_local_model(loader, 2, learning_rate=1e-3, optimizer_type="adam", save_checkpoints=False)
    g_loss, g_acc, g_ep, g_n = gold["metrics"]
    assert (m.epochs_completed, m.samples_processed) == (int(g_ep), int(g_n))
    assert abs(m.loss - g_loss) < 1e-4 and abs(m.accuracy - g_acc) < 1e-9 and m.training_time > 0
    wts = model.get_model_weights()
    assert all(v.is_cuda for v in wts.values())
    adam_trajectory_check(gold, "w", wts, OM.init_weights(MODEL, 12), 1e-3)
    ev = trainer.evaluate_model(loader)
    assert ev["total_samples"] == 40 and 0.0 <= ev["overall_accuracy"] <= 1.0
    logits = model.eval()(x.to(cuda_device))
    ref = OM.forward(MODEL, {k: v.cpu() for k, v in wts.items()}, x, train=False)
    np.testing.assert_allclose(logits.cpu().numpy(), ref.numpy(), rtol=1e-4, atol=1e-5)
    assert abs(ev["overall_accuracy"] - float((ref.argmax(1) == y).float().mean())) < 1e-9
    with pytest.raises(TrainingError, match="Local training failed: Unknown optimizer type: lion"):
        trainer.train_local_model(loader, 1, optimizer_type="lion")


def test_dropout_mask_injected_and_philox_rate(cuda_device):
    x, y = _data(9, 32)
    w = OM.init_weights(MODEL, 3)
    eng = _engine(cuda_device, 1, 32, dropout_rate=0.25)
    eng.set_client_weights(0, w)
    eng.load_data([x], [y])
    keep = (torch.rand(1, 32, 128, generator=torch.Generator().manual_seed(1)) >= 0.25)
    eng.drop_keep = keep.to(torch.uint8).to(cuda_device).contiguous()
    eng.forward_backward()
    loss, logits, grads = OT.loss_and_grads(MODEL, w, x, y, train=True, dropout_rate=0.25, masks=[keep[0].float()])
    np.testing.assert_allclose(eng.ws_array("logits", torch.float32, 10)[0].cpu().numpy(), logits.numpy(), rtol=1e-4, atol=1e-5)
    got = eng.layout.views(eng.G[0])
    for name, g in grads.items():
        np.testing.assert_allclose(got[name].cpu().numpy(), g.numpy(), rtol=2e-3, atol=2e-5 * float(g.abs().max()), err_msg=name)
    # Philox-generated mask: keeps ~75 % of the active units, scaled by 1/(1-p)
    eng.drop_keep = None
    eng.forward_backward()
    h = eng.ws_array("h", torch.float32, 128)[0].clone()
    eng.dropout_rate = 0.0
    eng.forward_backward()
    h0 = eng.ws_array("h", torch.float32, 128)[0].clone()
    active = h0 > 0
    kept = (h > 0) & active
    rate = kept.sum().item() / active.sum().item()
    assert 0.70 < rate < 0.80
    torch.testing.assert_close(h[kept], h0[kept] / 0.75, rtol=1e-4, atol=1e-6)     # split-K atomics reorder fc1's sum


def test_per_sample_dp_step_vs_oracle(cuda_device):
    """North-star kernel (2): per-sample clip + noise.  Oracle = vmap(grad) restatement (parity unpinned upstream)."""
    B = 16
    x, y = _data(77, B)
    w = {k: v * 3.0 for k, v in OM.init_weights(MODEL, 8).items()}       # larger weights -> norms straddle C
    eng = _engine(cuda_device, 2, B)
    lay = eng.layout
    eps, dlt = 1.0, 1e-5
    C = float(OD.per_sample_norms(OD.per_sample_grads(MODEL, w, x, y)).median())      # norms straddle C
    sigma = OPV.gaussian_sigma(C, eps, dlt)
    zrows = torch.randn((2, lay.ld), generator=torch.Generator().manual_seed(4))
    eng.configure_dp("per_sample", C, sigma, zrows.to(cuda_device))
    for k in range(2):
        eng.set_client_weights(k, w)
    eng.load_data([x, x[:9]], [y, y[:9]])
    eng.forward_backward()
    for k, n in enumerate([B, 9]):
        z = {name: zrows[k, lay.offsets[name]:lay.offsets[name] + v.numel()].view(v.shape) for name, v in w.items()}
        gbar, norms, s = OD.dp_sgd_grad(MODEL, w, x[:n], y[:n], C, eps, dlt, z=None)     # z=None: no noise term
        got_norm = eng.ws_array("norm2", torch.float32, 1)[k, :n, 0].sqrt().cpu()
        np.testing.assert_allclose(got_norm.numpy(), norms.numpy(), rtol=2e-5)
        assert (norms > C).any() and (norms < C).any() if k == 0 else True
        # G holds sum_i clip(g_i); the optimizer kernel adds sigma*z and divides by B
        gsum = lay.views(eng.G[k])
        for name in w:
            ref = gbar[name] * n
            np.testing.assert_allclose(gsum[name].cpu().numpy(), ref.numpy(), rtol=2e-3, atol=3e-5 * float(ref.abs().max()) + 1e-7, err_msg=name)
    # one SGD step (momentum buffer = grad at t = 1) exposes (sum + sigma z)/B through the weight update
    eng.train(1, 0.5, "sgd")
    z = {name: zrows[0, lay.offsets[name]:lay.offsets[name] + v.numel()].view(v.shape) for name, v in w.items()}
    gbar, _, _ = OD.dp_sgd_grad(MODEL, w, x, y, C, eps, dlt, z=z)
    got = eng.client_weights(0, "cpu")
    for name in w:
        ref = w[name] - 0.5 * gbar[name]
        # the update is dominated by 0.5 * sigma * z / B (sigma = 4.84 C): compare on that scale
        np.testing.assert_allclose(got[name].numpy(), ref.numpy(), rtol=1e-5, atol=1e-6 * sigma, err_msg=name)


def test_per_sample_dp_philox_noise_statistics(cuda_device):
    """With C tiny every gradient is clipped to ~0, so one SGD step moves the weights by lr * sigma * z / B."""
    B = 32
    x, y = _data(5, B)
    w = OM.init_weights(MODEL, 2)
    eng = _engine(cuda_device, 1, B)
    eng.configure_dp("per_sample", 1e-12, 2.0)
    eng.set_client_weights(0, w)
    eng.load_data([x], [y])
    w0 = eng.W[0, :eng.layout.P].clone()
    eng.train(1, 1.0, "sgd")
    z = (w0 - eng.W[0, :eng.layout.P]) * B / 2.0
    assert abs(z.mean().item()) < 5e-3 and abs(z.std().item() - 1) < 5e-3
    assert 0.5 <= z.abs().mean().item() <= 2.0            # the reference's window (privacy_validator.py:104-108)
