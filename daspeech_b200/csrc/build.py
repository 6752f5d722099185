"""Build libdagb200.so (in-tree) with nvcc for sm_100a.  Called by __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["cabi.cu", "lsg.cu", "dag_dp.cu", "dag_prep.cu", "dag_dp4.cu", "dag_viterbi3.cu", "dag_grad.cu", "dag_grad4.cu", "dag_posterior.cu", "dag_glat.cu", "dag_decode.cu", "dag_links.cu", "xchg.cu"]
OUT = os.path.join(HERE, "libdagb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v", "-ccbin", "g++"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    hdrs = [os.path.join(HERE, "common.cuh"), os.path.join(HERE, "..", "..", "include", "dagb200.h"),
            os.path.abspath(__file__)]
    hdrs += [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".cuh")]
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(HERE, src)
        o = os.path.join(HERE, src[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [NVCC] + FLAGS + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append("== %s\n%s" % (src, out))
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on %s" % src)
    with open(os.path.join(HERE, "build.log"), "a" if not force else "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    if procs or _stale(OUT, objs):
        subprocess.check_call([NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "g++"])
    return OUT


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
