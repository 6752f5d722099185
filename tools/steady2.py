import ctypes, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
ops = importlib.import_module("daspeech_b200.custom_ops.dag_loss")
k = ops.get_dag_kernel(); lib = k.lib
dev = torch.device("cuda", 0)
def prof():
    buf = (ctypes.c_float * 5)(); lib.dagb200_get_profile(ctypes.cast(buf, ctypes.c_void_p), 5); return [round(x, 3) for x in buf]
def run(tag, seed, nvml=False):
    if nvml:
        import pynvml; pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0); print("sm clock", pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
    match, links, olen, tlen, go = bench.make_inputs(torch, dev, 64, 1024, 256, 1023, 4096, seed)
    for _ in range(3):
        a, b = k.dag_loss(match, links, olen, tlen, True, 1); gm, gl = k.dag_loss_backward(go, a, b, match, links, olen, tlen, 2, 2)
    torch.cuda.synchronize()
    lib.dagb200_set_profile(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(20):
        a, b = k.dag_loss(match, links, olen, tlen, True, 1)
        gm, gl = k.dag_loss_backward(go, a, b, match, links, olen, tlen, 2, 2)
    e1.record(); torch.cuda.synchronize()
    print(tag, "seed", seed, "ms/step %.3f" % (e0.elapsed_time(e1) / 20), prof(), "match ptr %x links ptr %x" % (match.data_ptr(), links.data_ptr()))
    lib.dagb200_set_profile(0)
run("a", 1); run("b", 1234); run("c", 1); run("d-nvml", 1234, True)
