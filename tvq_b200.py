"""Import shim: `import tvq_b200` loads the package that lives in `t-vq-vae-trajgen_b200/`
(a directory name Python cannot import directly because of the hyphens)."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "t-vq-vae-trajgen_b200")
_spec = importlib.util.spec_from_file_location(
    "tvq_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_module = importlib.util.module_from_spec(_spec)
sys.modules["tvq_b200"] = _module
_spec.loader.exec_module(_module)
