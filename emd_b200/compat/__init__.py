"""Import-name shims: make ``import gsplat`` / ``import diff_gauss`` resolve to emd_b200.

Either put this directory on ``PYTHONPATH`` (``PYTHONPATH=/path/to/emd_b200/compat``)
or call ``emd_b200.compat.install()`` before the reference code is imported.
"""
import importlib
import os
import sys


def install(force: bool = False) -> None:
    here = os.path.dirname(os.path.abspath(__file__))
    for name in ("gsplat", "diff_gauss"):
        if name in sys.modules and not force:
            mod = sys.modules[name]
            if getattr(mod, "__emd_b200__", False):
                continue
            raise RuntimeError(f"emd_b200.compat.install(): a different `{name}` is already imported")
    if here not in sys.path:
        sys.path.insert(0, here)
    for name in ("gsplat", "gsplat.rendering", "gsplat.cuda", "gsplat.cuda._wrapper", "diff_gauss"):
        importlib.import_module(name)
