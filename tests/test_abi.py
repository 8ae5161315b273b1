"""The C-ABI library loads without a GPU and exports every symbol that include/mdhs_b200.h declares, with the
argument lists the ctypes binding expects (no compute calls here)."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _header_prototypes():
    hdr = open(os.path.join(ROOT, "include", "mdhs_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|int64_t)\s+(mdhs_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S):
        args = [a.strip() for a in m.group(2).replace("\n", " ").split(",") if a.strip() and a.strip() != "void"]
        sig = ""
        for a in args:
            if "*" in a:
                sig += "p"
            elif a.startswith("int64_t"):
                sig += "l"
            elif a.startswith("uint64_t"):
                sig += "u"
            elif a.startswith("float"):
                sig += "f"
            else:
                sig += "i"
        protos[m.group(1)] = sig
    return protos


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    import mdhs_b200  # noqa: F401
    from mdhs_b200 import _lib
    lib = _lib.lib()
    protos = _header_prototypes()
    assert len(protos) >= 35
    for name, sig in protos.items():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        if name in ("mdhs_abi_version", "mdhs_launch_count"):
            continue
        assert _lib.SIGNATURES.get(name) == sig, (name, _lib.SIGNATURES.get(name), sig)
    for name in _lib.SIGNATURES:
        assert name in protos, f"{name} bound in _lib.py but missing from the header"
    assert lib.mdhs_abi_version() == 4


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "multimodal-diagnosis-ham-spine_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_cpu_tensors_are_rejected_loudly():
    import pytest
    import torch
    import mdhs_b200  # noqa: F401
    from mdhs_b200 import _lib, ops
    a = torch.zeros(128, 64, dtype=torch.bfloat16)
    with pytest.raises(_lib.MdhsError):
        ops.gemm(a, a)
