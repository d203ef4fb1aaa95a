"""The C-ABI library loads without a GPU, exports every symbol include/rabe_b200.h declares, and
fails loudly (RB_ECUDA, never a CPU fallback) when no device is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "rabe_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from rabe_b200 import _lib
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), n
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)
    assert L.rb_strerror(0) == b"ok" and b"CUDA" in L.rb_strerror(-4)


def test_product_never_imports_oracle():
    for dp, _, files in os.walk(os.path.join(ROOT, "rabe_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "oracle/" not in src and "liboracle" not in src, f


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rabe_b200 import _lib
    from rabe_b200.engine import Engine
    with pytest.raises(_lib.RabeB200Error) as ei:
        Engine(0)
    assert ei.value.status == _lib.RB_ECUDA
