"""Per-node cost of a dependent chain of tiny library kernels, eager vs inside one CUDA graph."""
import sys, time
sys.path.insert(0, ".")
import torch
import ehgr_b200
from ehgr_b200 import _lib

C = 96
dev = torch.device("cuda")
stats = torch.rand(2 * C, dtype=torch.float64, device=dev) + 1
g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
out = torch.empty((4, C), device=dev)
N = 1000


def chain():
    sp = _lib.stream_ptr(dev)
    for _ in range(N):
        _lib.call("ehgr_bn_finalize", stats.data_ptr(), 1000, g.data_ptr(), b.data_ptr(), rm.data_ptr(), rv.data_ptr(), 0.1, 1e-5, 1,
                  out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(), C, sp)


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / N


chain()
print("eager   us/kernel", round(timed(chain), 2))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        chain()
gr.replay()
print("graph   us/kernel", round(timed(gr.replay), 2))
