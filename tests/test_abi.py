"""The C-ABI shared library loads here (no GPU) and exports every symbol include/*.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    from garlic_b200 import build
    return build.build()


def declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for fn in os.listdir(inc):
        if fn.endswith(".h"):
            txt = open(os.path.join(inc, fn)).read()
            txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
            names |= set(re.findall(r"\b(garlic_gpu_\w+)\s*\(", txt))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    so = _ensure_built()
    lib = ctypes.CDLL(so)
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "missing export " + s


def test_python_binding_lists_the_same_symbols():
    from garlic_b200 import api
    assert sorted(api.EXPORTS) == declared_symbols()


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from garlic_b200 import api
    _ensure_built()
    with pytest.raises(api.GarlicError):
        api.GarlicGPU(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "garlic_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, fn)).read()
                assert "oracle" not in txt.replace("the oracle", "").replace("literal oracle", "") or fn in ("synth.py",), fn
