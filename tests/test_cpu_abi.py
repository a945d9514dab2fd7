"""CPU: the C-ABI library loads, exports every symbol include/emd_b200.h declares, the header matches
the sources, and the ctypes signature table agrees with the prototypes (argument counts).  No compute."""
import ctypes
import re
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))


def _header_protos():
    text = (ROOT / "include" / "emd_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"^([\w\s\*]+?\b(emd_\w+)\s*\(([^;]*?)\))\s*;", text, re.M | re.S):
        args = m.group(3).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",")])
        protos[m.group(2)] = n
    return protos


def test_header_is_current():
    import gen_header
    assert gen_header.render() == (ROOT / "include" / "emd_b200.h").read_text(), \
        "include/emd_b200.h is stale: run python tools/gen_header.py --write"


def test_library_exports_every_declared_symbol():
    from emd_b200 import build
    lib = build.build()
    L = ctypes.CDLL(str(lib))
    protos = _header_protos()
    assert len(protos) >= 30
    for name in protos:
        assert hasattr(L, name), f"{lib} does not export {name}"
    out = subprocess.run(["nm", "-D", "--defined-only", str(lib)], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln and ln.split()[-1].startswith("emd_")}
    assert exported == set(protos), exported ^ set(protos)


def test_ctypes_table_matches_header():
    from emd_b200 import _C
    protos = _header_protos()
    assert set(_C._SIGS) == set(protos), set(_C._SIGS) ^ set(protos)
    for name, (_res, args) in _C._SIGS.items():
        assert len(args) == protos[name], f"{name}: ctypes has {len(args)} args, header has {protos[name]}"


def test_no_cpu_fallback():
    """Ops refuse CPU tensors instead of silently computing somewhere else."""
    import torch
    import emd_b200
    from emd_b200 import _C
    with pytest.raises(_C.EmdError):
        emd_b200.spherical_harmonics(1, torch.randn(4, 3), torch.randn(4, 4, 3))
    z = torch.zeros(4, 3)
    with pytest.raises(_C.EmdError):
        emd_b200.rasterization(z, torch.zeros(4, 4), z, torch.zeros(4), z, torch.eye(4)[None], torch.eye(3)[None],
                               32, 32, packed=False)


def test_experimental_tensor_core_dense_path_is_off_by_default(monkeypatch):
    """csrc/deform_net_tc.cu has not run on hardware yet: nothing may select it implicitly."""
    import subprocess
    import sys
    code = ("import os; os.environ.pop('EMD_DENSE_TC', None); from emd_b200 import _C; "
            "print(_C.lib().emd_dense_tc_enabled())")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip().endswith("0")
