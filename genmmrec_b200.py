"""Import shim: `import genmmrec_b200` loads the package that lives in the directory
`generative-multimodal-recommendation_b200/` (a name Python cannot import directly because of
the hyphens).  The shim replaces itself in `sys.modules` with the real package, so
`genmmrec_b200.models.diffmm`, `genmmrec_b200.common.trainer` ... resolve as usual.
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "generative-multimodal-recommendation_b200")
_spec = importlib.util.spec_from_file_location(
    "genmmrec_b200", os.path.join(_PKG_DIR, "__init__.py"),
    submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["genmmrec_b200"] = _mod
_spec.loader.exec_module(_mod)
