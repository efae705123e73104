"""Stand-alone K1 (temporal shift) and K2-K6 (ACTION) kernels at the MobileNetV2 insertion sites,
BASELINE config #2 size (B=32 clips -> 256 frames): CUDA-event time and algorithmic GB/s against the
measured HBM peak.  This is also the command ncu captures for profiles/ (run with --reps 1 under ncu).

    python tools/bench_shift_action.py [--frames 256] [--reps 10] [--only shift,action]
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ehgr_b200 as E
from ehgr_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAK = 6451.5
try:
    PEAK = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

SITES = [(24, 56), (32, 28), (64, 14), (96, 14), (160, 7)]       # (C, H) — SURVEY §8a A1


def flush_l2(buf):
    buf.zero_()


def time_call(fn, reps, flush):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush_l2(flush)
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
    ap.add_argument("--only", default="shift,action")
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    only = set(args.only.split(","))
    nt, T = args.frames, 8
    dev = torch.device("cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    rows = []

    if "shift" in only:
        for c, h in SITES:
            for dt in (torch.float32, torch.bfloat16):
                for layout in ("nchw", "nhwc"):
                    x = torch.randn(nt, c, h, h, device=dev).to(dt)
                    if layout == "nhwc":
                        x = x.contiguous(memory_format=torch.channels_last)
                    nbytes = 2 * x.numel() * x.element_size()
                    for direction, bwd in (("fwd", False), ("bwd", True)):
                        secs = time_call(lambda: E.temporal_shift_module._run_shift(x, T, c // 8, bwd), args.reps, flush)
                        rows.append({"kernel": f"temporal_shift_{direction}", "site": f"C{c}@{h}", "dtype": str(dt)[6:],
                                     "layout": layout, "us": round(secs * 1e6, 2), "GBps": round(nbytes / secs / 1e9, 1),
                                     "frac_of_hbm_peak": round(nbytes / secs / 1e9 / PEAK, 4), "MB": round(nbytes / 1e6, 2)})
                        print(rows[-1], flush=True)

    if "action" in only:
        for c, h in SITES:
            with contextlib.redirect_stdout(io.StringIO()):
                mod = E.Action(torch.nn.Conv2d(c, 6 * c, 1, bias=False), n_segment=T, shift_div=8).to(dev)
            for dt in (torch.bfloat16,):
                x = torch.randn(nt, c, h, h, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
                x.requires_grad_(True)
                gy = torch.randn(nt, c, h, h, device=dev).to(dt).contiguous(memory_format=torch.channels_last)

                def fwd_bwd():
                    y = E.action_ops.gated(mod, x, dt)
                    y.backward(gy)
                fwd_bwd()
                torch.cuda.synchronize()
                acc = {}
                for _ in range(args.reps):
                    flush_l2(flush)
                    rec = _lib.KernelTimer.begin()
                    fwd_bwd()
                    out = _lib.KernelTimer.end(rec)
                    for k, v in out.items():
                        a = acc.setdefault(k, {"ms": [], "bytes": v["bytes"], "launches": v["launches"]})
                        a["ms"].append(v["ms"])
                for k, a in acc.items():
                    ms = sorted(a["ms"])[len(a["ms"]) // 2]
                    gbs = a["bytes"] / (ms / 1e3) / 1e9 if a["bytes"] else None
                    rows.append({"kernel": k, "site": f"C{c}@{h}", "dtype": str(dt)[6:], "launches": a["launches"],
                                 "us": round(ms * 1e3, 2), "GBps": None if gbs is None else round(gbs, 1),
                                 "frac_of_hbm_peak": None if gbs is None else round(gbs / PEAK, 4),
                                 "MB": round(a["bytes"] / 1e6, 2)})
                    print(rows[-1], flush=True)
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"frames": nt, "peak_gbs": PEAK, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
