"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden [--ref /root/reference]

The reference ships no fixtures (SURVEY.md "Five facts" #4); these files pin the oracle to the
reference's own behaviour on seeded inputs.  /root/reference does not exist on the GPU box, so
only the vectors travel.  Large tensors are stored as strided samples (every STRIDE-th element)
plus per-tensor float64 sums and L2 norms; inputs are regenerated from seeds and guarded by
checksums.
"""
from __future__ import annotations

import argparse
import os
import sys
import types
from datetime import datetime

import numpy as np
import torch

STRIDE = 97
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def digest(prefix: str, w: dict, out: dict) -> None:
    for k, t in w.items():
        a = t.detach().reshape(-1).to(torch.float32).numpy()
        out[f"{prefix}/{k}/sample"] = a[::STRIDE].copy() if a.size > 4096 else a.copy()
        out[f"{prefix}/{k}/sum"] = np.float64(a.astype(np.float64).sum())
        out[f"{prefix}/{k}/l2"] = np.float64(np.sqrt((a.astype(np.float64) ** 2).sum()))


def seeded_weights(model, seed):
    from oracle import models as OM
    return OM.init_weights(model, seed)


def seeded_batches(model, seed, n, bs):
    from oracle import models as OM
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((n,) + OM.input_shape(model), generator=g)
    y = torch.randint(0, 10, (n,), generator=g)
    return x, y, [(x[i:i + bs], y[i:i + bs]) for i in range(0, n, bs)]


def convergence_sequence(seed: int = 21, rounds: int = 8):
    """Seeded sequence of global models with shrinking steps (SimpleCNN shapes): [(weights, accuracy_metrics)]."""
    from oracle import models as OM
    g = torch.Generator().manual_seed(seed)
    w = OM.init_weights("simple_cnn", seed)
    seq = []
    for r in range(rounds):
        w = {k: v + (0.05 * 0.35 ** r) * torch.randn(v.shape, generator=g) for k, v in w.items()}
        seq.append((w, {"test_accuracy": min(0.8, 0.5 + 0.08 * r) - (0.03 if r == 5 else 0.0), "train_loss": 1.5 * 0.7 ** r + (0.2 if r == 6 else 0.0)}))
    return seq


def convergence_golden(out_dir: str) -> None:
    """---- 8. ConvergenceDetector: the unmodified reference on the seeded sequence ----"""
    from src.aggregation.convergence import create_convergence_detector
    from src.shared.models import GlobalModel
    out = {}
    for kind in ("standard", "adaptive"):
        det = create_convergence_detector(kind, patience=3)
        prev, rows, stops = None, [], []
        for r, (w, accm) in enumerate(convergence_sequence()):
            cur = GlobalModel(round_number=r, model_weights=w, accuracy_metrics=accm, participating_clients=["c0"],
                              convergence_score=0.0, created_at=datetime.now())
            m = det.calculate_convergence_metrics(cur, prev)
            rows.append([m.weight_change_norm, m.relative_weight_change, m.accuracy_change, m.loss_change, m.convergence_score,
                         float(m.is_converged), m.confidence, det.convergence_threshold])
            stops.append(det.should_stop_early()[1])
            prev = cur
        out[f"{kind}/rows"] = np.asarray(rows, dtype=np.float64)
        out[f"{kind}/stop_reasons"] = np.asarray(stops)
        out[f"{kind}/trend"] = np.asarray(det.get_convergence_summary()["recent_performance"]["convergence_trend"])
    np.savez_compressed(os.path.join(out_dir, "convergence.npz"), **out)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--only-convergence", action="store_true", help="write tests/golden/convergence.npz only")
    args = ap.parse_args()
    if args.only_convergence:
        sys.dont_write_bytecode = True
        sys.path.insert(0, args.ref)
        torch.set_num_threads(1)
        convergence_golden(OUT)
        print("convergence golden written to", OUT)
        return
    sys.dont_write_bytecode = True
    sys.path.insert(0, args.ref)
    sys.modules.setdefault("lz4", types.ModuleType("lz4"))
    sys.modules.setdefault("lz4.frame", types.ModuleType("lz4.frame"))
    sys.modules["lz4"].frame = sys.modules["lz4.frame"]
    torch.set_num_threads(1)

    from src.aggregation.fedavg import FedAvgAggregator
    from src.shared.compression import QuantizationCompressor, TopKSparsificationCompressor
    from src.shared.models import ModelUpdate
    from src.shared.models_pytorch import ModelFactory
    from src.shared.privacy import DifferentialPrivacyEngine, GradientClipper, create_privacy_engine
    from src.shared.training import LocalTrainer

    os.makedirs(OUT, exist_ok=True)
    meta = {"torch": torch.__version__, "generated": datetime.now().isoformat(timespec="seconds")}

    # ---- 1. forward + one-step gradients --------------------------------------------------
    for model, n in (("simple_cnn", 6), ("cifar10_cnn", 6)):
        out = {}
        w0 = seeded_weights(model, 11)
        x, y, _ = seeded_batches(model, 21, n, n)
        m = ModelFactory.create_model(model, dropout_rate=0.0)
        m.set_model_weights(w0)
        m.train()
        logits = m(x)
        loss = torch.nn.CrossEntropyLoss()(logits, y)
        loss.backward()
        out["x_sum"] = np.float64(x.double().sum())
        out["y"] = y.numpy()
        out["logits"] = logits.detach().numpy()
        out["loss"] = np.float64(loss.item())
        digest("grad", {k: p.grad for k, p in m.named_parameters()}, out)
        m.eval()
        out["logits_eval"] = m(x).detach().numpy()
        np.savez_compressed(os.path.join(OUT, f"forward_{model}.npz"), **out)

    # ---- 2. LocalTrainer.train_local_model ------------------------------------------------
    for model, n, bs in (("simple_cnn", 40, 8), ("cifar10_cnn", 24, 8)):
        for opt in ("adam", "sgd", "adamw"):
            out = {}
            w0 = seeded_weights(model, 12)
            x, y, batches = seeded_batches(model, 22, n, bs)
            m = ModelFactory.create_model(model, dropout_rate=0.0)
            m.set_model_weights(w0)
            tr = LocalTrainer(m, device=torch.device("cpu"))
            met = tr.train_local_model(batches, epochs=2, learning_rate=1e-3 if opt != "sgd" else 1e-2,
                                       optimizer_type=opt, save_checkpoints=False)
            out["x_sum"] = np.float64(x.double().sum())
            out["metrics"] = np.array([met.loss, met.accuracy, met.epochs_completed, met.samples_processed],
                                      dtype=np.float64)
            digest("w", m.get_model_weights(), out)
            if model == "cifar10_cnn":
                digest("buf", {k: b.float() for k, b in m.named_buffers() if "num_batches" not in k}, out)
            np.savez_compressed(os.path.join(OUT, f"train_{model}_{opt}.npz"), **out)

    # ---- 3. update-level DP: clip, sigma, injected noise ------------------------------------
    class InjectedNoise:
        """stand-in for engine.noise_generator (privacy.py:275): noise = sigma * z with given z."""
        def __init__(self, z):
            self.z, self.calls = z, []

        def add_noise_to_gradients(self, g, sens, eps, delta):
            import math
            sigma = sens * math.sqrt(2 * math.log(1.25 / delta)) / eps
            self.calls.append((sens, sigma))
            return {k: t + sigma * self.z[k] for k, t in g.items()}

    out = {}
    gen = torch.Generator().manual_seed(31)
    shapes = {"a.weight": (17, 5, 3, 3), "a.bias": (17,), "b.weight": (33, 129), "b.bias": (33,)}
    for tag, scale in (("big", 0.05), ("small", 0.001)):
        g = {k: torch.randn(s, generator=gen) * scale for k, s in shapes.items()}
        z = {k: torch.randn(s, generator=gen) for k, s in shapes.items()}
        clipped, norm = GradientClipper(1.0).clip_gradients(g)
        eng = create_privacy_engine(epsilon=1.0, delta=1e-5, max_grad_norm=1.0)
        eng.noise_generator = InjectedNoise(z)
        noisy = eng.add_noise(g, 1.0, 1e-5)
        for k in shapes:
            out[f"{tag}/g/{k}"] = g[k].numpy()
            out[f"{tag}/z/{k}"] = z[k].numpy()
            out[f"{tag}/clipped/{k}"] = clipped[k].numpy()
            out[f"{tag}/noisy/{k}"] = noisy[k].numpy()
        out[f"{tag}/norm"] = np.float64(norm)
        out[f"{tag}/sens_sigma"] = np.array(eng.noise_generator.calls[0], dtype=np.float64)
        rem = eng.budget_tracker.get_remaining_budget()
        out[f"{tag}/remaining"] = np.array(rem, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "privacy_update_level.npz"), **out)

    # ---- 4. FedAvg --------------------------------------------------------------------------
    out = {}
    gen = torch.Generator().manual_seed(41)
    shapes = {"l1.weight": (8, 3, 3, 3), "l1.bias": (8,), "l2.weight": (10, 131), "l2.bias": (10,)}
    K = 7
    ns = [int(v) for v in torch.randint(100, 1000, (K,), generator=gen)]
    ns[3] = ns[5]                                            # a tie for the max_clients truncation
    losses = [float(v) for v in torch.rand(K, generator=gen) * 2]
    ups = []
    for i in range(K):
        w = {k: torch.randn(s, generator=gen) * 0.1 for k, s in shapes.items()}
        ups.append(ModelUpdate(f"client_{i}", 3, w, ns[i], losses[i], 0.1, 0.8, datetime.now()))
        for k in shapes:
            out[f"theta/{i}/{k}"] = w[k].numpy()
    out["num_samples"] = np.array(ns)
    out["losses"] = np.array(losses, dtype=np.float64)
    gm = FedAvgAggregator(validate_updates=False).aggregate_updates(ups)
    for k in shapes:
        out[f"by_samples/{k}"] = gm.model_weights[k].numpy()
    cw = [0.5, 1.5, 0.0, 2.0, 1.0, 0.25, 3.0]
    out["custom_weights"] = np.array(cw)
    gm2 = FedAvgAggregator(validate_updates=False).aggregate_updates(ups, cw)
    for k in shapes:
        out[f"custom/{k}"] = gm2.model_weights[k].numpy()
    agg3 = FedAvgAggregator(min_clients=2, max_clients=4, validate_updates=False)
    gm3 = agg3.aggregate_updates(ups)
    out["top4_participants"] = np.array([int(c.split("_")[1]) for c in gm3.participating_clients])
    for k in shapes:
        out[f"top4/{k}"] = gm3.model_weights[k].numpy()
    out["avg_loss"] = np.float64(agg3.aggregation_history[-1]["avg_training_loss"])
    np.savez_compressed(os.path.join(OUT, "fedavg.npz"), **out)

    # ---- 5. codecs ----------------------------------------------------------------------------
    out = {}
    gen = torch.Generator().manual_seed(51)
    x = torch.randn(4099, generator=gen)
    x[7] = 0.0
    out["x"] = x.numpy()
    for bits, sym in ((8, True), (4, True), (16, True), (8, False)):
        qc = QuantizationCompressor(bits=bits, symmetric=sym)
        q, scale, zp = qc._quantize_tensor(x)
        dq = qc._dequantize_tensor(q, scale, zp, x.shape, str(x.dtype))
        tag = f"q{bits}{'s' if sym else 'a'}"
        out[f"{tag}/q"] = q.numpy()
        out[f"{tag}/scale_zp"] = np.array([scale, zp], dtype=np.float64)
        out[f"{tag}/dq"] = dq.numpy()
    for sp in (0.9, 0.5, 0.99995):
        tk = TopKSparsificationCompressor(sparsity_ratio=sp)
        vals, idx, shp = tk._sparsify_tensor(x.reshape(4099))
        dense = tk._desparsify_tensor(vals, idx, shp, str(x.dtype))
        out[f"topk{sp}/idx"] = idx.numpy()
        out[f"topk{sp}/vals"] = vals.numpy()
        out[f"topk{sp}/dense"] = dense.numpy()
    np.savez_compressed(os.path.join(OUT, "codecs.npz"), **out)

    # ---- 6. a whole round: 3 clients, train + injected-noise DP + FedAvg --------------------
    from oracle import round as OR
    out = {}
    model = "simple_cnn"
    w0 = seeded_weights(model, 13)
    gen = torch.Generator().manual_seed(61)
    ups = []
    for c in range(3):
        x, y = OR.synthetic_client_data(model, c, n=64 + 32 * c)
        m = ModelFactory.create_model(model, dropout_rate=0.0)
        m.set_model_weights(w0)
        tr = LocalTrainer(m, device=torch.device("cpu"))
        met = tr.train_local_model([(x[i:i + 32], y[i:i + 32]) for i in range(0, x.shape[0], 32)], epochs=1,
                                   save_checkpoints=False)
        cur = m.get_model_weights()
        delta = {k: cur[k] - w0[k] for k in cur}            # federated_trainer.py:437-443
        z = {k: torch.randn(v.shape, generator=gen) for k, v in cur.items()}
        eng = create_privacy_engine(1.0, 1e-5, 1.0)
        eng.noise_generator = InjectedNoise({k: v * 1e-3 for k, v in z.items()})
        noisy = eng.add_noise(delta, 1.0, 1e-5)
        m.set_model_weights({k: w0[k] + noisy[k] for k in w0})   # :454-462
        out[f"client{c}/metrics"] = np.array([met.loss, met.accuracy, met.samples_processed], dtype=np.float64)
        out[f"client{c}/sens_sigma"] = np.array(eng.noise_generator.calls[0], dtype=np.float64)
        ups.append(ModelUpdate(f"client_{c}", 0, m.get_model_weights(), met.samples_processed, met.loss,
                               1.0, 0.8, datetime.now()))
    gm = FedAvgAggregator(validate_updates=False).aggregate_updates(ups)
    digest("global", gm.model_weights, out)
    np.savez_compressed(os.path.join(OUT, "round_simple_cnn.npz"), **out)

    # ---- 7. DataPartitioner (data side) ----------------------------------------
    import random
    from src.shared.data_loader import DataPartitioner

    class _Labelled(torch.utils.data.Dataset):
        def __init__(self, y):
            self.y = y

        def __len__(self):
            return len(self.y)

        def __getitem__(self, i):
            return torch.zeros(1), int(self.y[i])

    out = {}
    y = np.random.default_rng(3).integers(0, 10, 3000)
    out["labels"] = y
    for strat in ("iid", "non_iid", "pathological"):
        for nc in (7, 20):
            random.seed(5)
            np.random.seed(6)
            part = DataPartitioner(_Labelled(y), nc, strat)
            for cid, idx in part.client_indices.items():
                out[f"{strat}/{nc}/{cid}"] = np.asarray(idx, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "partition.npz"), **out)

    convergence_golden(OUT)

    with open(os.path.join(OUT, "README.md"), "w") as f:
        f.write("# Golden vectors\n\nWritten by `python -m oracle.make_golden` from the unmodified reference "
                f"(torch {meta['torch']}, 1 CPU thread, {meta['generated']}).\n"
                "Inputs are regenerated from seeds by the tests; big tensors are stored as every "
                f"{STRIDE}th element plus float64 sum / L2 norm.\n")
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
