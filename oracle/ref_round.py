"""One federated round driven through the reference's OWN classes (test / benchmark infrastructure, never the product).

`oracle/_ref/` holds byte-identical copies of the reference's hot-path modules (oracle/build_ref.py).  The only code
restated here is the client glue that cannot be imported (src/client/federated_trainer.py fails with NameError at :262
and needs lz4): `_download_global_model` (:367-388), `_perform_local_training` (:390-426), `_apply_differential_privacy`
(:428-469) and `_upload_model_update` (:471-500), in that order, for every client sequentially, followed by the
coordinator's `FedAvgAggregator.aggregate_updates` (src/aggregation/fedavg.py:56-124)."""
from __future__ import annotations

import os
import sys
import time
from datetime import datetime
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import build_ref

_mods = None


def load():
    """Import the staged reference modules (as the package `src`, exactly as they import each other)."""
    global _mods
    if _mods is None:
        if not build_ref.available():
            raise ImportError("oracle/_ref is not staged: run `python -m oracle.build_ref` in the build container")
        sys.dont_write_bytecode = True
        if build_ref.DEST not in sys.path:
            sys.path.insert(0, build_ref.DEST)
        from src.aggregation.fedavg import FedAvgAggregator
        from src.shared.models import ModelUpdate
        from src.shared.models_pytorch import ModelFactory
        from src.shared.privacy import create_privacy_engine
        from src.shared.training import LocalTrainer
        _mods = dict(FedAvgAggregator=FedAvgAggregator, ModelUpdate=ModelUpdate, ModelFactory=ModelFactory,
                     create_privacy_engine=create_privacy_engine, LocalTrainer=LocalTrainer)
    return _mods


def batches(x: torch.Tensor, y: torch.Tensor, bs: int) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """What an unshuffled DataLoader(batch_size=bs) yields (src/shared/data_loader.py:356-362), pre-batched."""
    return [(x[i:i + bs], y[i:i + bs]) for i in range(0, x.shape[0], bs)]


def federated_round(model_name: str, global_weights: Dict[str, torch.Tensor], data: Sequence[Tuple[torch.Tensor, torch.Tensor]],
                    dp: bool = True, epsilon: float = 1.0, delta: float = 1e-5, max_grad_norm: float = 1.0,
                    batch_size: int = 32, lr: float = 1e-3, optimizer: str = "adam", epochs: int = 1,
                    dropout_rate: Optional[float] = None, round_number: int = 1):
    """Returns (aggregated weights, info).  `dropout_rate=None` keeps the model's own default (0.25 / 0.3), i.e. the
    configuration the B200 arm of bench.py runs."""
    R = load()
    kw = {} if dropout_rate is None else {"dropout_rate": dropout_rate}
    updates, t_train, t_dp = [], 0.0, 0.0
    for cid, (x, y) in enumerate(data):
        model = R["ModelFactory"].create_model(model_name, **kw)
        model.set_model_weights(global_weights)                                   # federated_trainer.py:378
        trainer = R["LocalTrainer"](model, device=torch.device("cpu"))
        m = trainer.train_local_model(batches(x, y, batch_size), epochs=epochs, learning_rate=lr, optimizer_type=optimizer,
                                      save_checkpoints=False)                     # :401-408
        t_train += m.training_time
        if dp:                                                                    # :428-469
            t0 = time.perf_counter()
            cur = model.get_model_weights()
            grads = {k: cur[k] - global_weights[k] for k in cur}
            engine = R["create_privacy_engine"](epsilon=epsilon, delta=delta, max_grad_norm=max_grad_norm)
            noisy = engine.add_noise(grads, epsilon=epsilon, delta=delta)
            model.set_model_weights({k: global_weights[k] + noisy[k] for k in global_weights})
            t_dp += time.perf_counter() - t0
        updates.append(R["ModelUpdate"](client_id=f"client-{cid}", round_number=round_number,
                                        model_weights=model.get_model_weights(), num_samples=m.samples_processed,
                                        training_loss=m.loss, privacy_budget_used=epsilon if dp else 0.0,
                                        compression_ratio=0.8, timestamp=datetime.now()))          # :476-485
    agg = R["FedAvgAggregator"](min_clients=1, validate_updates=False)
    t0 = time.perf_counter()
    gm = agg.aggregate_updates(updates)
    t_agg = time.perf_counter() - t0
    return gm.model_weights, {"num_samples": [u.num_samples for u in updates], "losses": [u.training_loss for u in updates],
                              "train_s": t_train, "dp_s": t_dp, "aggregate_s": t_agg}
