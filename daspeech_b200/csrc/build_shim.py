"""Build dag_loss_fn_b200.so (in-tree): the pybind module with the reference's four native entry points
(DASpeech/custom_ops/dag_loss.cpp:19-29) on top of libdagb200.so.  Plain g++; called by __graft_entry__.build()."""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "dag_loss_fn_shim.cpp")
OUT = os.path.join(HERE, "dag_loss_fn_b200.so")


def build(force=False):
    lib = os.path.join(HERE, "libdagb200.so")
    deps = [SRC, os.path.join(HERE, "..", "..", "include", "dagb200.h"), os.path.abspath(__file__)]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    if not os.path.exists(lib):
        raise RuntimeError("build libdagb200.so first (daspeech_b200/csrc/build.py)")
    import torch
    from torch.utils import cpp_extension as ce
    inc = ce.include_paths() + [sysconfig.get_paths()["include"], "/usr/local/cuda/include"]
    libdirs = ce.library_paths()
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", SRC, "-o", OUT, "-DTORCH_EXTENSION_NAME=dag_loss_fn_b200",
           "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    cmd += ["-I" + i for i in inc]
    cmd += ["-L" + d for d in libdirs] + ["-L" + HERE]
    cmd += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-ldagb200", "-Wl,-rpath,$ORIGIN"]
    cmd += ["-Wl,-rpath," + d for d in libdirs]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        sys.stderr.write(p.stdout)
        raise RuntimeError("g++ failed on dag_loss_fn_shim.cpp")
    return OUT


def load():
    """Import the built module (needs torch imported first so that libtorch symbols resolve)."""
    import importlib.util
    import torch  # noqa: F401
    if not os.path.exists(OUT):
        raise RuntimeError("%s is missing: run daspeech_b200/csrc/build_shim.py" % OUT)
    spec = importlib.util.spec_from_file_location("dag_loss_fn_b200", OUT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="-f" in sys.argv))
