"""CPU: libdagb200.so builds, loads, exports every symbol include/dagb200.h declares, and rejects bad
arguments on the host side (no kernel is launched, so no GPU is needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dagb200.h")


@pytest.fixture(scope="module")
def lib():
    from daspeech_b200.csrc import build
    build.build()
    from daspeech_b200 import _lib
    return _lib.load()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dagb200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_four_reference_entry_points():
    syms = declared_symbols()
    for name in ("dagb200_dag_loss", "dagb200_dag_loss_backward", "dagb200_dag_best_alignment",
                 "dagb200_logsoftmax_gather"):
        assert name in syms


def test_every_declared_symbol_is_exported_and_bound(lib):
    from daspeech_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 8
    for name in syms:
        assert hasattr(lib, name), name
        assert name in _lib.SIGNATURES, "ctypes binding missing for " + name
    assert lib.dagb200_version() == 100


def test_argument_errors_are_reported_without_touching_the_gpu(lib):
    # bad config (reference: TORCH_CHECK "config should be 1~4", dag_loss.cu:351)
    rc = lib.dagb200_dag_loss(1, 1, 1, 1, 1, 1, 0, 2, 4, 8, 7, 1, 9, None, 0, None, None)
    assert rc == -1 and b"config should be 1~4" in lib.dagb200_last_error()
    # unsupported lattice dtype (reference dispatch covers float/double only)
    rc = lib.dagb200_dag_loss(1, 1, 1, 1, 1, 1, 1, 2, 4, 8, 7, 1, 1, None, 0, None, None)
    assert rc == -2 and b"float32 or float64" in lib.dagb200_last_error()
    # null pointers
    rc = lib.dagb200_dag_loss(None, None, None, None, None, None, 0, 2, 4, 8, 7, 1, 1, None, 0, None, None)
    assert rc == -1
    rc = lib.dagb200_dag_loss_backward(1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 2, 4, 8, 7, 3, 1, None)
    assert rc == -1 and b"config1" in lib.dagb200_last_error()
    rc = lib.dagb200_dag_best_alignment(1, 1, 1, 1, None, 1, 0, 2, 4, 8, 7, 1, None, 0, None, None)
    assert rc == -4
    rc = lib.dagb200_logsoftmax_gather(1, 9, 1, 0, 0, 1, 1, 1, 1, 1, 2, 4, 8, 3, 0, None)
    assert rc == -2
    assert lib.dagb200_best_alignment_workspace_bytes(2, 4, 8, 7) >= 2 * 4 * 8 * 6
    # empty batch is a no-op
    assert lib.dagb200_dag_loss(None, None, None, None, None, None, 0, 0, 4, 8, 7, 1, 1, None, 0, None, None) == 0
    assert lib.dagb200_dag_loss_workspace_bytes(64, 256, 1024, 1023) > 64 * 4 * 1024 * 1024


def test_missing_library_fails_loudly(monkeypatch):
    from daspeech_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdagb200.so")
    with pytest.raises(RuntimeError, match="no CPU/torch fallback"):
        _lib.load()


def test_pybind_shim_exports_the_reference_entry_points():
    """daspeech_b200/csrc/dag_loss_fn_b200.so: the module a maintainer returns from get_dag_kernel() (INTEGRATION.md option B)
    must expose exactly the four callables of DASpeech/custom_ops/dag_loss.cpp:24-29 (import only; no GPU here)."""
    from daspeech_b200.csrc import build_shim
    build_shim.build()
    mod = build_shim.load()
    names = sorted(n for n in dir(mod) if not n.startswith("_"))
    assert names == ["dag_best_alignment", "dag_loss", "dag_loss_backward", "logsoftmax_gather"]
    for n in names:
        assert callable(getattr(mod, n))
