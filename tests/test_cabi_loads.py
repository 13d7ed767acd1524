"""CPU: the C-ABI library loads and exports every symbol include/tfel_cuda.h declares; without a CUDA device
every entry point fails loudly (no CPU fallback); the product never imports the oracle."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "tfel_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tfelcu_[a-z0-9_]+)\s*\(", src)))


def test_binding_lists_every_header_symbol():
    from tfel_b200 import capi
    assert sorted(capi.SYMBOLS) == header_symbols()


def test_library_exports_every_symbol():
    from tfel_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tfel_b200", "csrc")])
    L = capi.lib()
    for name in header_symbols():
        assert hasattr(L, name), name
    assert b"sm_100a" in L.tfelcu_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from tfel_b200 import capi
    with pytest.raises(capi.TfelCudaError, match="no CUDA device"):
        capi.Context(0)


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "tfel_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "oracle" not in text.replace("is never", "").lower() or f in ("__init__.py", "capi.py"), (base, f)
    code = "import sys; import tfel_b200; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
