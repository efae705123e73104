"""Per-kernel micro-benchmark on the MobileNetV2 layer shapes at BASELINE config #2 size
(B=32 clips -> 256 frames, bf16): CUDA-event time, achieved algorithmic GB/s vs the measured HBM peak.

    python tools/bench_kernels.py [--frames 256] [--only dw_fwd,pw_fwd,...] [--reps 20]

Used to find which kernel to optimise and as the command ncu captures (profiles/).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ehgr_b200 as E
from ehgr_b200 import _lib

f = E.fused
PEAK = 6451.5
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

# (name, cin, hidden, cout, H_in, stride) per InvertedResidual of MobileNetV2 at 224x224 input
BLOCKS = [("f2", 16, 96, 24, 112, 2), ("f3", 24, 144, 24, 56, 1), ("f4", 24, 144, 32, 56, 2), ("f5", 32, 192, 32, 28, 1),
          ("f7", 32, 192, 64, 28, 2), ("f8", 64, 384, 64, 14, 1), ("f11", 64, 384, 96, 14, 1), ("f12", 96, 576, 96, 14, 1),
          ("f14", 96, 576, 160, 14, 2), ("f15", 160, 960, 160, 7, 1), ("f17", 160, 960, 320, 7, 1)]


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(reps):
        flush.zero_()                           # evict L2 between iterations
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
    ap.add_argument("--only", default="")
    ap.add_argument("--blocks", default="")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--engine", type=int, default=0)
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    blocks = set(filter(None, args.blocks.split(",")))
    nt = args.frames
    dt = torch.bfloat16
    sp = _lib.stream_ptr(torch.device("cuda"))
    rows = []

    def vec(c, lo=0.5, hi=1.5):
        return torch.rand(c, device="cuda") * (hi - lo) + lo

    def report(kind, name, shape, secs, nbytes, flops=0):
        gbs = nbytes / secs / 1e9
        rows.append((kind, name, shape, secs * 1e6, gbs, gbs / PEAK, flops / secs / 1e12))
        print(f"{kind:10s} {name:5s} {shape:28s} {secs*1e6:9.1f} us  {gbs:8.1f} GB/s  {gbs/PEAK*100:5.1f}% of HBM peak  {flops/secs/1e12:7.2f} TFLOP/s",
              flush=True)

    for name, cin, hid, cout, h, stride in BLOCKS:
        if blocks and name not in blocks:
            continue
        ho = (h - 1) // stride + 1
        m_in, m_out = nt * h * h, nt * ho * ho
        x = torch.randn(m_in, cin, device="cuda").to(dt)
        raw1 = torch.empty(m_in, hid, device="cuda", dtype=dt)
        raw2 = torch.empty(m_out, hid, device="cuda", dtype=dt)
        raw3 = torch.empty(m_out, cout, device="cuda", dtype=dt)
        g1 = torch.randn(m_in, hid, device="cuda").to(dt)
        g2 = torch.randn(m_out, hid, device="cuda").to(dt)
        g3 = torch.randn(m_out, cout, device="cuda").to(dt)
        w1 = torch.randn(hid, cin, device="cuda") * 0.2
        w2 = torch.randn(hid, 9, device="cuda") * 0.3
        w3 = torch.randn(cout, hid, device="cuda") * 0.05
        w1h, w3h = w1.to(dt), w3.to(dt)     # bf16 mirrors, as the fused chain passes them
        s1, b1, s2, b2, s3, b3 = vec(hid), vec(hid, -.2, .2), vec(hid), vec(hid, -.2, .2), vec(cout), vec(cout, -.2, .2)
        ca1, cb1, cc1 = vec(hid), vec(hid, -.1, .1), vec(hid, -.1, .1)
        ca2, cb2, cc2 = vec(hid), vec(hid, -.1, .1), vec(hid, -.1, .1)
        ca3, cb3, cc3 = vec(cout), vec(cout, -.1, .1), vec(cout, -.1, .1)
        st1 = torch.zeros(2 * hid, dtype=torch.float64, device="cuda")
        st3 = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
        dwg = torch.zeros(hid * 9, dtype=torch.float32, device="cuda")
        dw1 = torch.zeros(hid * cin, dtype=torch.float32, device="cuda")
        dw3 = torch.zeros(hid * cout, dtype=torch.float32, device="cuda")
        gx = torch.empty(m_in, cin, device="cuda", dtype=dt)
        d1 = torch.randn(m_in, hid, device="cuda").to(dt)
        d3 = torch.randn(m_out, cout, device="cuda").to(dt)
        es = 2
        tests = {
            "pw_fwd": (lambda: _lib.call("ehgr_pw_gemm_w16", ctypes.byref(f.op_plain(x)), w1.data_ptr(), w1h.data_ptr(), 0, raw1.data_ptr(), 0, st1.data_ptr(),
                                         m_in, cin, hid, 1, args.engine, sp), f"expand {cin}->{hid} @{h}", m_in * (cin + hid) * es, 2 * m_in * cin * hid),
            "dw_fwd": (lambda: _lib.call("ehgr_dw_fwd", ctypes.byref(f.op_affine(raw1, s1, b1, True)), w2.data_ptr(), raw2.data_ptr(),
                                         st1.data_ptr(), nt, h, h, hid, stride, 1, sp), f"dw {hid} @{h} s{stride}", (m_in + m_out) * hid * es, 18 * m_out * hid),
            "pw_proj": (lambda: _lib.call("ehgr_pw_gemm_w16", ctypes.byref(f.op_affine(raw2, s2, b2, True)), w3.data_ptr(), w3h.data_ptr(), 0, raw3.data_ptr(), 0,
                                          st3.data_ptr(), m_out, hid, cout, 1, args.engine, sp), f"project {hid}->{cout} @{ho}", m_out * (hid + cout) * es,
                        2 * m_out * hid * cout),
            # backward of a pointwise layer as the chain issues it: d(raw) materialised once (row_apply),
            # then dgrad and wgrad read it as a PLAIN operand
            "draw3": (lambda: _lib.call("ehgr_row_apply", ctypes.byref(f.op_bnbwd(g3, raw3, ca3, cb3, cc3, s3, b3, False)), 0,
                                        d3.data_ptr(), m_out, cout, 1, sp), f"d(raw) {cout} @{ho}", 3 * m_out * cout * es, 0),
            "pw_dgrad3": (lambda: _lib.call("ehgr_pw_gemm_w16", ctypes.byref(f.op_plain(d3)), w3.data_ptr(), w3h.data_ptr(), 1,
                                            g2.data_ptr(), 0, 0, m_out, cout, hid, 1, args.engine, sp), f"dgrad {cout}->{hid} @{ho}",
                          m_out * (cout + hid) * es, 2 * m_out * hid * cout),
            "pw_wgrad3": (lambda: _lib.call("ehgr_pw_wgrad", ctypes.byref(f.op_plain(d3)),
                                            ctypes.byref(f.op_affine(raw2, s2, b2, True)), dw3.data_ptr(), m_out, hid, cout, 1, args.engine, sp),
                          f"wgrad {cout}x{hid} @{ho}", m_out * (cout + hid) * es, 2 * m_out * hid * cout),
            "bn_reduce": (lambda: _lib.call("ehgr_bn_bwd_reduce", g2.data_ptr(), raw2.data_ptr(), s2.data_ptr(), b2.data_ptr(), 1,
                                            st1.data_ptr(), m_out, hid, 1, sp), f"bn-bwd reduce {hid} @{ho}", 2 * m_out * hid * es, 0),
            "dw_dgrad": (lambda: _lib.call("ehgr_dw_dgrad", ctypes.byref(f.op_bnbwd(g2, raw2, ca2, cb2, cc2, s2, b2, True)), w2.data_ptr(),
                                           g1.data_ptr(), nt, h, h, hid, stride, 1, sp), f"dw dgrad {hid} @{h} s{stride}",
                         (2 * m_out + m_in) * hid * es, 18 * m_in * hid),
            "dw_wgrad": (lambda: _lib.call("ehgr_dw_wgrad", ctypes.byref(f.op_bnbwd(g2, raw2, ca2, cb2, cc2, s2, b2, True)),
                                           ctypes.byref(f.op_affine(raw1, s1, b1, True)), dwg.data_ptr(), nt, h, h, hid, stride, 1, sp),
                         f"dw wgrad {hid} @{h} s{stride}", (2 * m_out + m_in) * hid * es, 18 * m_out * hid),
            "dw_bwd": (lambda: _lib.call("ehgr_dw_bwd", ctypes.byref(f.op_bnbwd(g2, raw2, ca2, cb2, cc2, s2, b2, True)),
                                         ctypes.byref(f.op_affine(raw1, s1, b1, True)), w2.data_ptr(), g1.data_ptr(), dwg.data_ptr(),
                                         nt, h, h, hid, stride, 1, sp),
                       f"dw fused bwd {hid} @{h} s{stride}", (2 * m_out + 2 * m_in) * hid * es, 18 * (m_in + m_out) * hid),
            "draw1": (lambda: _lib.call("ehgr_row_apply", ctypes.byref(f.op_bnbwd(g1, raw1, ca1, cb1, cc1, s1, b1, True)), 0,
                                        d1.data_ptr(), m_in, hid, 1, sp), f"d(raw) {hid} @{h}", 3 * m_in * hid * es, 0),
            "pw_dgrad1": (lambda: _lib.call("ehgr_pw_gemm_w16", ctypes.byref(f.op_plain(d1)), w1.data_ptr(), w1h.data_ptr(), 1,
                                            gx.data_ptr(), 0, 0, m_in, hid, cin, 1, args.engine, sp), f"dgrad {hid}->{cin} @{h}",
                          m_in * (hid + cin) * es, 2 * m_in * hid * cin),
            "pw_wgrad1": (lambda: _lib.call("ehgr_pw_wgrad", ctypes.byref(f.op_plain(d1)),
                                            ctypes.byref(f.op_plain(x)), dw1.data_ptr(), m_in, cin, hid, 1, args.engine, sp),
                          f"wgrad {hid}x{cin} @{h}", m_in * (hid + cin) * es, 2 * m_in * hid * cin),
            "row_apply": (lambda: _lib.call("ehgr_row_apply", ctypes.byref(f.op_affine(raw3, s3, b3, False)), g3.data_ptr(), raw3.data_ptr(),
                                            m_out, cout, 1, sp), f"bn-apply+res {cout} @{ho}", 3 * m_out * cout * es, 0),
        }
        for kind, (fn, shape, nbytes, flops) in tests.items():
            if only and kind not in only:
                continue
            report(kind, name, shape, timeit(fn, args.reps), nbytes, flops)
    tot = {}
    for kind, _, _, us, _, _, _ in rows:
        tot[kind] = tot.get(kind, 0.0) + us
    print("sum over the listed blocks (us):", {k: round(v, 1) for k, v in tot.items()})


if __name__ == "__main__":
    main()
