"""Build recipe for oracle/_ref/dag_loss_fn*.so  (TEST INFRASTRUCTURE, never imported by the product).

Compiles the UNMODIFIED reference CUDA extension from the sources where they lie
(/root/reference/DASpeech/custom_ops/{dag_loss.cpp,dag_loss.cu,dag_best_alignment.cu,
logsoftmax_gather.cu}) for sm_100a with the flags the reference itself uses
(dag_loss.py:51-62: -O3 -DOF_SOFTMAX_USE_FAST_MATH).  Only the built shared object is
written, into oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).  No
reference source is copied into this repository.

On the GPU box /root/reference does not exist: there `load_ref()` only dlopens the
prebuilt .so (or returns None) and the differential tests skip when it is absent.
"""
import glob
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_OPS = "/root/reference/DASpeech/custom_ops"
NAME = "dag_loss_fn"


def build(verbose=False):
    if not os.path.isdir(REF_OPS):
        return None
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils.cpp_extension import load
    srcs = [os.path.join(REF_OPS, f) for f in
            ("dag_loss.cpp", "dag_loss.cu", "dag_best_alignment.cu", "logsoftmax_gather.cu")]
    load(NAME, sources=srcs,
         extra_cflags=["-DOF_SOFTMAX_USE_FAST_MATH", "-O3"],
         extra_cuda_cflags=["-DOF_SOFTMAX_USE_FAST_MATH", "-O3"],
         build_directory=OUT, verbose=verbose, is_python_module=False)
    # keep only the shared object (drop ninja files / objects that embed source paths)
    for f in (os.path.join(OUT, n) for n in os.listdir(OUT)):
        if not f.endswith(".so"):
            try:
                os.remove(f)
            except OSError:
                pass
    return so_path()


def so_path():
    c = sorted(glob.glob(os.path.join(OUT, NAME + "*.so")))
    return c[0] if c else None


def load_ref():
    """Import the prebuilt reference extension as a python module (needs torch + a GPU to run)."""
    p = so_path()
    if p is None:
        return None
    import torch  # noqa: F401  (registers libtorch symbols before dlopen)
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print("reference extension:", p)
