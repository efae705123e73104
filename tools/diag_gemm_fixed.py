"""Fixed cost vs per-tile cost of the tcgen05 GEMM: time vs M for a few (K, N)."""
import sys, ctypes
sys.path.insert(0, ".")
import torch
import ehgr_b200 as E
from ehgr_b200 import _lib
f = E.fused
dt = torch.bfloat16
sp = _lib.stream_ptr(torch.device("cuda"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


DBG = int(sys.argv[1]) if len(sys.argv) > 1 else 0
if DBG:
    _lib.lib().ehgr_debug_set(DBG)
print("dbg", DBG)
for K, N in ((384, 64), (960, 160)):
    w = torch.randn(N, K, device="cuda") * 0.05
    wh = w.to(dt)
    for tiles in (1, 148):
        M = 128 * tiles
        a = torch.randn(M, K, device="cuda").to(dt)
        out = torch.empty(M, N, device="cuda", dtype=dt)
        fn = lambda: _lib.call("ehgr_pw_gemm_w16", ctypes.byref(f.op_plain(a)), w.data_ptr(), wh.data_ptr(), 0, out.data_ptr(), 0, 0,
                               M, K, N, 1, 2, sp)
        print(f"K={K:4d} N={N:4d} tiles={tiles:5d} M={M:7d}: {timeit(fn):7.1f} us", flush=True)
