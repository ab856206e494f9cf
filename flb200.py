"""Import shim: the product package lives in the directory the contract names
(``federated-learning-for-privacy-preserving-image-classification_b200/``), whose name is not a
valid Python identifier.  ``import flb200`` loads that directory as the package ``flb200``."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "federated-learning-for-privacy-preserving-image-classification_b200")
_spec = importlib.util.spec_from_file_location(
    "flb200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["flb200"] = _mod
_spec.loader.exec_module(_mod)
