"""Build an instrumented / A-B variant of libdagb200.so into tools/_dbg/ (not the product):
   python tools/build_variant.py <name> <file.cu> [-DFLAG ...]   -> tools/_dbg/libdagb200_<name>.so
Run with DAGB200_LIB=tools/_dbg/libdagb200_<name>.so."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from daspeech_b200.csrc import build as b
name, src = sys.argv[1], sys.argv[2]
flags = sys.argv[3:]
out_dir = os.path.join(ROOT, "tools", "_dbg")
os.makedirs(out_dir, exist_ok=True)
obj = os.path.join(out_dir, src[:-3] + "_" + name + ".o")
subprocess.check_call([b.NVCC] + [f for f in b.FLAGS if f not in ("-Xptxas", "-v")] + flags + ["-c", os.path.join(b.HERE, src), "-o", obj])
objs = [obj if s == src else os.path.join(b.HERE, s[:-3] + ".o") for s in b.SOURCES]
so = os.path.join(out_dir, "libdagb200_%s.so" % name)
subprocess.check_call([b.NVCC, "-shared", "-o", so] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "g++"])
print(so)
