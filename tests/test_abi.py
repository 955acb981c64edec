"""The C-ABI shared library loads and exports every symbol include/b200fock.h declares
(no compute calls: runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "b200fock.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_abi():
    names = _declared()
    assert "b200_apply_gate1" in names and "b200_gather_reduce" in names
    assert len(names) >= 20


def test_library_exports_every_declared_symbol():
    from strawberryfields_b200 import build, lib

    if not os.path.exists(build.LIB) and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        pytest.skip("no prebuilt library and no nvcc")
    build.build()
    handle = ctypes.CDLL(build.LIB)
    for name in _declared():
        assert hasattr(handle, name), name
    # the Python binding covers the same set
    assert sorted(lib.SIGNATURES) == _declared()


def test_no_cpu_fallback_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from strawberryfields_b200 import B200FockBackend, lib

    with pytest.raises(lib.B200Error):
        B200FockBackend().begin_circuit(2, cutoff_dim=4)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "strawberryfields_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
