"""CPU: the C-ABI library loads and exports exactly what include/flb.h declares (no compute calls)."""
import os
import re

import pytest

import flb200
from flb200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "flb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.findall(r"\b(?:int|long long|const char\*)\s+(flb_\w+)\s*\(", src)


def test_library_exports_every_header_symbol():
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/flb.h but missing from libflb.so"


def test_python_binding_matches_header():
    declared = set(_header_functions()) - {"flb_last_error"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_header_argument_counts_match_binding():
    src = open(os.path.join(ROOT, "include", "flb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for name, args in re.findall(r"\b(?:int|long long)\s+(flb_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        n = 0 if args.strip() in ("", "void") else len(args.split(","))
        assert n == len(_lib.SIGNATURES[name]), name


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from flb200 import ops
    with pytest.raises(_lib.FlbError):
        ops.fedavg_weighted_sum(torch.zeros(2, 32), [0.5, 0.5])
    from flb200.privacy import PrivacyError, create_privacy_engine
    eng = create_privacy_engine()
    with pytest.raises(PrivacyError):
        eng.add_noise({"w": torch.ones(4)}, 1.0, 1e-5)
    assert _lib.load().flb_version() == 100


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "federated-learning-for-privacy-preserving-image-classification_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
