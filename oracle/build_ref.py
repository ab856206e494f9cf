"""Stage the UNMODIFIED reference modules of the hot path under oracle/_ref/ (test / benchmark infrastructure).

    python -m oracle.build_ref [--ref /root/reference]

The reference is pure Python: there is nothing to compile, and its top-level packages do not import as a whole
(SyntaxError at src/coordinator/grpc_server.py:582, NameError at src/client/federated_trainer.py:262, missing lz4), so
`pip install` is not an option.  The nine leaf modules the path consists of DO import on their own; this recipe copies
them byte for byte (sha256 recorded in oracle/_ref/MANIFEST.json) so that the benchmark's CPU arm can drive the reference's
own `LocalTrainer` / `DifferentialPrivacyEngine` / `FedAvgAggregator` on the GPU box, where /root/reference does not exist.
oracle/_ref/ is git-ignored (the sources never enter this repository's history) but travels with the tree to the GPU box,
like the built libflb.so.  Nothing under the product package imports it; only bench.py's CPU legs and tests do."""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
MODULES = [
    "src/__init__.py",
    "src/shared/__init__.py",
    "src/shared/models.py",             # value types (ModelUpdate, GlobalModel, PrivacyConfig, TrainingMetrics)
    "src/shared/interfaces.py",
    "src/shared/models_pytorch.py",     # SimpleCNN / CIFAR10CNN / ModelFactory
    "src/shared/training.py",           # LocalTrainer
    "src/shared/privacy.py",            # DifferentialPrivacyEngine
    "src/shared/validation.py",         # imported by fedavg.py
    "src/aggregation/__init__.py",
    "src/aggregation/fedavg.py",        # FedAvgAggregator
]


def build(ref: str = "/root/reference", dest: str = DEST) -> str:
    if not os.path.isdir(os.path.join(ref, "src")):
        raise FileNotFoundError(f"{ref}/src not found (the reference only exists in the build container)")
    manifest = {}
    for rel in MODULES:
        src, dst = os.path.join(ref, rel), os.path.join(dest, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": ref, "files": manifest}, f, indent=1)
    return dest


def available(dest: str = DEST) -> bool:
    return all(os.path.exists(os.path.join(dest, rel)) for rel in MODULES)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    print(build(ap.parse_args().ref))
