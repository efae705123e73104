import sys, ctypes
sys.path.insert(0, ".")
import torch
import ehgr_b200 as E
from ehgr_b200 import _lib
nt, h, w, cout = 3, 32, 32, 32
x = torch.randn(nt, 3, h, w, device="cuda")
wt = torch.randn(cout, 27, device="cuda")
ho = wo = 16
out = torch.empty(nt * ho * wo, cout, device="cuda")
stats = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
_lib.call("ehgr_stem_fwd", x.data_ptr(), wt.data_ptr(), out.data_ptr(), stats.data_ptr(), nt, h, w, cout, 0, 0, _lib.stream_ptr(x.device))
torch.cuda.synchronize()
print("ok", out.abs().sum().item())
