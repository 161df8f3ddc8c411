"""Import the UNMODIFIED reference backend (TEST / BENCH INFRASTRUCTURE, never the product path).

Two places can hold it: ``baseline/_ref/torchPIV`` (pip-installed by ``baseline/install_ref.sh``;
git-ignored, travels to the GPU box with the working tree) and ``/root/reference/src/torchPIV`` (build
container only).  Users: ``tests/golden/make_golden.py`` (fixture generation), CPU tests that skip themselves
when the reference is absent, and ``bench.py`` (``--impl reference`` and the torch-CUDA comparator leg).
The reference package imports its Qt GUI at package import and ``imageio`` in the backend (unused); four
empty stub modules are enough to load ``PIVbackend.py`` as is (recipe: SURVEY.md section 8c).
"""
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = [os.environ.get("TORCHPIV_REF"), os.path.join(_HERE, "baseline", "_ref", "torchPIV"),
               "/root/reference/src/torchPIV"]
REF_ROOT = next((c for c in _CANDIDATES if c and os.path.isfile(os.path.join(c, "PIVbackend.py"))),
                _CANDIDATES[-1])


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "PIVbackend.py"))


def load_ref(root: str = REF_ROOT):
    for name in ("imageio", "imageio.v3", "PyQt5", "PyQt5.QtWidgets"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["imageio"].v3 = sys.modules["imageio.v3"]
    sys.modules["PyQt5.QtWidgets"].QMessageBox = object
    pkg = types.ModuleType("torchPIV")
    pkg.__path__ = [root]
    sys.modules["torchPIV"] = pkg

    def _load(name):
        spec = importlib.util.spec_from_file_location(f"torchPIV.{name}", f"{root}/{name}.py")
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"torchPIV.{name}"] = mod
        spec.loader.exec_module(mod)
        return mod

    _load("PlotterFunctions")
    return _load("PIVbackend")
