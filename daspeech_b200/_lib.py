"""ctypes loader for libdagb200.so (the C-ABI declared in include/dagb200.h).

There is NO fallback: if the shared library is missing or a call fails the caller gets a RuntimeError.
Build it with `python -c "import __graft_entry__ as g; g.build()"` or `python daspeech_b200/csrc/build.py`.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# DAGB200_LIB: an alternative build of the same library (instrumented / A-B variants), never a fallback
LIB_PATH = os.environ.get("DAGB200_LIB") or os.path.join(_HERE, "csrc", "libdagb200.so")

_lock = threading.Lock()
_lib = None

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_sz = ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/dagb200.h declares
SIGNATURES = {
    "dagb200_version": (_int, []),
    "dagb200_last_error": (ctypes.c_char_p, []),
    "dagb200_set_exact": (None, [_int]),
    "dagb200_get_exact": (_int, []),
    "dagb200_set_profile": (None, [_int]),
    "dagb200_get_profile": (_int, [_vp, _int]),
    "dagb200_logsoftmax_gather": (_int, [_vp, _int, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64,
                                         _int, _int, _int, _int, _int, _vp]),
    "dagb200_logsoftmax_gather_argmax": (_int, [_vp, _int, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _vp,
                                                _int, _int, _int, _int, _int, _vp]),
    "dagb200_logsoftmax_gather_backward": (_int, [_vp, _int, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64,
                                                  _int, _int, _int, _int, _vp]),
    "dagb200_dag_loss_workspace_bytes": (_sz, [_int, _int, _int, _int]),
    "dagb200_dag_loss": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _int,
                                _vp, _sz, _vp, _vp]),
    "dagb200_dag_loss_backward": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int,
                                         _int, _int, _int, _int, _int, _int, _vp]),
    "dagb200_dag_loss_backward_workspace_bytes": (_sz, [_int, _int, _int, _int]),
    "dagb200_dag_loss_backward_ws": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int,
                                            _int, _int, _int, _int, _int, _int, _vp, _sz, _vp]),
    "dagb200_dag_posterior": (_int, [_vp, _vp, _vp, _int, _int, _int, _int, _vp]),
    "dagb200_glat_force_emit": (_int, [_vp, _vp, _vp, _vp, _int, _int, _int, _int, _vp]),
    "dagb200_glat_alignment": (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _int, _vp]),
    "dagb200_decode_lookahead": (_int, [_vp, _vp, _vp, _vp, ctypes.c_float, _i64, _int, _int, _int, _vp, _vp, _vp, _vp]),
    "dagb200_decode_viterbi_finish": (_int, [_vp, _vp, _vp, _vp, ctypes.c_float, _i64, _int, _int, _int, _int, _vp, _vp, _vp,
                                             _vp, _vp]),
    "dagb200_best_alignment_workspace_bytes": (_sz, [_int, _int, _int, _int]),
    "dagb200_dag_best_alignment": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _int,
                                          _vp, _sz, _vp, _vp]),
    "dagb200_extract_links_workspace_bytes": (_sz, [_int, _int, _int, _int]),
    "dagb200_extract_links": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _vp, _sz, _vp]),
    "dagb200_peer_alloc": (_int, [_sz, ctypes.POINTER(ctypes.c_void_p)]),
    "dagb200_peer_free": (_int, [_vp]),
    "dagb200_peer_export": (_int, [_vp, _vp]),
    "dagb200_peer_open": (_int, [_vp, ctypes.POINTER(ctypes.c_void_p)]),
    "dagb200_peer_close": (_int, [_vp]),
    "dagb200_grad_exchange_slice": (_sz, [_sz, _int]),
    "dagb200_grad_exchange_create": (_int, [_vp, _vp, _vp, _sz, _int, _int, ctypes.POINTER(ctypes.c_void_p)]),
    "dagb200_grad_exchange": (_int, [_vp, _vp]),
    "dagb200_grad_exchange_status": (_int, [_vp, ctypes.POINTER(_int)]),
    "dagb200_grad_exchange_phases": (_int, [_vp, _vp]),
    "dagb200_grad_exchange_destroy": (_int, [_vp]),
    "dagb200_grad_exchange_nvls": (_int, [_vp, _sz, _int, _int, _int, _vp]),
}


def load():
    """dlopen libdagb200.so and bind every entry point.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "daspeech_b200: %s is missing -- the CUDA extension has not been built and there is no "
                "CPU/torch fallback. Run `python daspeech_b200/csrc/build.py`." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().dagb200_last_error()
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else ""))
