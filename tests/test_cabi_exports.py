"""CPU-side checks of the boundary: the shared library loads, exports every symbol include/distaff_gpu.h declares, and
fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "distaff_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from distaff_b200 import backend
    if not os.path.exists(backend.LIB_PATH):
        pytest.skip("libdistaff_gpu.so not built (run __graft_entry__.build())")
    L = ctypes.CDLL(backend.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), n
    bound = set(backend.EXPORTS) | set(backend.VOID_EXPORTS) | {"dg_last_error"}
    assert set(names) == bound, set(names) ^ bound


def test_no_cpu_fallback_without_a_device():
    import numpy as np
    from distaff_b200 import backend
    if not os.path.exists(backend.LIB_PATH):
        pytest.skip("libdistaff_gpu.so not built")
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    import distaff_b200
    with pytest.raises(backend.DgError) as e:
        distaff_b200.ntt(np.zeros((8, 2), dtype=np.uint64))
    assert e.value.code == -3 and "no CPU path" in str(e.value)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "distaff_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) and "build" not in dirpath:
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pyoracle" not in text and "liboracle" not in text and '"../../oracle' not in text and "oracle/" not in text.replace("the oracle", ""), f
