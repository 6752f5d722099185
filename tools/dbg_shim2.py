import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import build_ref
from daspeech_b200.csrc import build_shim
print("dlopenflags", sys.getdlopenflags())
ref = build_ref.load_ref()
print("ref", ref, ref.dag_loss.__doc__)
mod = build_shim.load()
print("shim", mod, mod.dag_loss.__doc__)
print(ref is mod, ref.dag_loss is mod.dag_loss)
