"""Import alias: the product package lives in the (non-identifier) directory
`robust-speech-emotion-recognition-via-dynamic-asymmetric-distillation-in-noisy-environments_b200/`;
`import dad_b200` loads that directory as the package `dad_b200` (sub-modules: dad_b200.model, ...)."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "robust-speech-emotion-recognition-via-dynamic-asymmetric-distillation-in-noisy-environments_b200")
_spec = importlib.util.spec_from_file_location("dad_b200", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["dad_b200"] = _mod
_spec.loader.exec_module(_mod)
