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


# ---- a host written in plain C against include/distaff_gpu.h (examples/prove_trace.c) ---------------------------------------------
def _build_c_host(tmp_path):
    import subprocess
    from distaff_b200 import backend
    if not os.path.exists(backend.LIB_PATH):
        pytest.skip("libdistaff_gpu.so not built")
    exe = str(tmp_path / "prove_trace")
    libdir = os.path.dirname(backend.LIB_PATH)
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "prove_trace.c"),
                           "-L", libdir, "-ldistaff_gpu", "-Wl,-rpath," + libdir, "-o", exe])
    return exe


def _write_trace(path, tr):
    import struct
    import numpy as np
    from distaff_b200 import felt
    with open(path, "wb") as f:
        f.write(struct.pack("<6IQ", tr.width, tr.ctx_depth, tr.loop_depth, len(tr.public_inputs), len(tr.outputs), 0, tr.length))
        f.write(felt.from_ints(tr.public_inputs).tobytes() if len(tr.public_inputs) else b"")
        f.write(felt.from_ints(tr.outputs).tobytes() if len(tr.outputs) else b"")
        f.write(np.ascontiguousarray(tr.registers).tobytes())


def test_c_host_compiles_against_the_header_and_fails_without_a_device(tmp_path):
    """the header is plain C (no C++ / torch types) and a C program links against the library; without a GPU dg_prove reports -3"""
    import subprocess
    from distaff_b200 import hostvm
    exe = _build_c_host(tmp_path)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present (covered by the gpu test below)")
    _write_trace(tmp_path / "trace.bin", hostvm.fibonacci(13))
    r = subprocess.run([exe, str(tmp_path / "trace.bin"), str(tmp_path / "proof.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU path" in r.stderr, (r.returncode, r.stderr)
    assert not os.path.exists(tmp_path / "proof.bin")


@pytest.mark.gpu
def test_c_host_proof_is_byte_identical_to_the_oracle(tmp_path):
    import subprocess
    from distaff_b200 import hostvm
    from oracle import pyoracle as po
    exe = _build_c_host(tmp_path)
    for tr, opts in ((hostvm.fibonacci(13), ()), (hostvm.collatz(27), ("16", "30", "8"))):
        _write_trace(tmp_path / "trace.bin", tr)
        r = subprocess.run([exe, str(tmp_path / "trace.bin"), str(tmp_path / "proof.bin"), *opts], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        kw = dict(ext=int(opts[0]), num_queries=int(opts[1]), grinding=int(opts[2])) if opts else {}
        ref = po.prove(tr.registers, tr.ctx_depth, tr.loop_depth, tr.public_inputs, tr.outputs, **kw)
        assert ref.error is None
        assert open(tmp_path / "proof.bin", "rb").read() == ref.proof
