"""Import the UNMODIFIED reference backend (TEST INFRASTRUCTURE, this container only).

``/root/reference`` exists only in the build container, never on the GPU box, so this is
used solely by ``tests/golden/make_golden.py`` (fixture generation) and by CPU tests that
skip themselves when the reference is absent.  The reference package imports its Qt GUI
at package import and ``imageio`` in the backend (unused); four empty stub modules are
enough to load ``PIVbackend.py`` as is (recipe: SURVEY.md section 8c).
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("TORCHPIV_REF", "/root/reference/src/torchPIV")


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
