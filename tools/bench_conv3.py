"""Per-layer timing of the depth decoder's implicit-GEMM convolutions at BASELINE config #2 size (256 frames, bf16).

    python tools/bench_conv3.py [--frames 256] [--reps 10]
"""
from __future__ import annotations

import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ehgr_b200 as E
from ehgr_b200 import _lib

f = E.fused
LAYERS = [(1280, 256, 7, 0), (256, 64, 7, 1), (64, 32, 14, 1), (32, 32, 28, 1)]   # cin, cout, stored grid, upsampled


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--layers", default="0,1,2,3", help="indices into LAYERS")
    a = ap.parse_args()
    nt = a.frames
    sp = _lib.stream_ptr(torch.device("cuda"))
    for cin, cout, hs, up in [LAYERS[int(i)] for i in a.layers.split(',')]:
        ho = hs << up
        M = nt * ho * ho
        x = torch.randn(nt, hs, hs, cin, device="cuda").to(torch.bfloat16)
        w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
        sc, sh = torch.rand(cin, device="cuda") + 0.5, torch.randn(cin, device="cuda") * 0.1
        wf = torch.empty(cout * 9 * cin, dtype=torch.bfloat16, device="cuda")
        wd = torch.empty_like(wf)
        _lib.call("ehgr_conv3_pack", w.data_ptr(), wf.data_ptr(), wd.data_ptr(), cout, cin, 1, sp)
        out = torch.empty(nt, ho, ho, cout, dtype=torch.bfloat16, device="cuda")
        gup = torch.empty(nt, ho, ho, cin, dtype=torch.bfloat16, device="cuda")
        dwp = torch.zeros(cout * 9 * cin, dtype=torch.float32, device="cuda")
        stats = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
        flops = 2.0 * M * 9 * cin * cout
        xu = torch.nn.functional.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=2, mode="nearest").permute(0, 2, 3, 1).contiguous().to(torch.bfloat16) if up else x
        variants = [("gather", f.op_conv3(x, None, None, 0, ho, ho, cin, up)), ("affine", f.op_conv3(x, sc, sh, 2, ho, ho, cin, up))]
        if up:       # what the fused chain runs: the upsample materialised, the operand gathered with TMA boxes
            variants.insert(0, ("tma", f.op_conv3(xu, None, None, 0, ho, ho, cin, 0)))
        for tag, op in variants:
            t = timeit(lambda: _lib.call("ehgr_pw_gemm_w16", ctypes.byref(op), w.data_ptr(), wf.data_ptr(), 0, out.data_ptr(), 0,
                                         stats.data_ptr(), M, 9 * cin, cout, 1, 2, sp), a.reps)
            print(f"fwd   {cin:5d}->{cout:4d} @{ho:3d} up={up} {tag:6s} {t * 1e6:9.1f} us  {flops / t / 1e12:7.1f} TFLOP/s")
            t = timeit(lambda: _lib.call("ehgr_pw_wgrad", ctypes.byref(f.op_plain(out)), ctypes.byref(op), dwp.data_ptr(), M, 9 * cin,
                                         cout, 1, 2, sp), a.reps)
            print(f"wgrad {cin:5d}x{cout:4d} @{ho:3d} up={up} {tag:6s} {t * 1e6:9.1f} us  {flops / t / 1e12:7.1f} TFLOP/s")
        t = timeit(lambda: _lib.call("ehgr_pw_gemm_w16", ctypes.byref(f.op_conv3(out, None, None, 0, ho, ho, cout, 0)), w.data_ptr(),
                                     wd.data_ptr(), 0, gup.data_ptr(), 0, 0, M, 9 * cout, cin, 1, 2, sp), a.reps)
        print(f"dgrad {cout:5d}->{cin:4d} @{ho:3d}              {t * 1e6:9.1f} us  {flops / t / 1e12:7.1f} TFLOP/s")


if __name__ == "__main__":
    main()
