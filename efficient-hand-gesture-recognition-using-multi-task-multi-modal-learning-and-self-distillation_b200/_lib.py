"""ctypes binding of libehgr_b200.so (C ABI: include/ehgr_b200.h).

There is NO CPU or PyTorch fallback: if the shared library cannot be loaded (or built with nvcc),
importing an operator that needs it raises.  Every wrapper takes torch CUDA tensors, passes raw
device pointers plus the current CUDA stream, and raises RuntimeError on a non-zero status.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_void_p
from pathlib import Path

import torch

_PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = _PKG_DIR / "csrc" / "libehgr_b200.so"

F32, BF16 = 0, 1
ABI_VERSION = 2
NCHW, NHWC = 0, 1

_lib = None


def _load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists() or os.environ.get("EHGR_REBUILD") == "1":
        from . import build as _build  # builds with nvcc; raises when nvcc is missing

        _build.build()
    lib = ctypes.CDLL(str(LIB_PATH))
    lib.ehgr_abi_version.restype = c_int
    lib.ehgr_status_string.restype = c_char_p
    lib.ehgr_status_string.argtypes = [c_int]
    lib.ehgr_launch_count.restype = c_longlong
    if lib.ehgr_abi_version() != ABI_VERSION:
        raise RuntimeError("libehgr_b200.so ABI version mismatch; rebuild with EHGR_REBUILD=1")
    _declare(lib)
    _lib = lib
    return lib


def lib() -> ctypes.CDLL:
    return _load()


def launch_count() -> int:
    return int(_load().ehgr_launch_count())


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = _load().ehgr_status_string(status).decode()
        raise RuntimeError(f"libehgr_b200 {what} failed: {msg} (status {status})")


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"libehgr_b200 supports float32 and bfloat16 activations, got {t.dtype}")


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "ehgr_b200 operators run on CUDA (sm_100a) only; there is no CPU fallback "
                f"(got a tensor on {t.device})")


def on_gpu(t) -> bool:
    """Whether `t` lives on a CUDA device — the one place the module wrappers ask before taking the library's path
    (CPU tensors run the plain modules: shape / policy tests only; the CPU test-suite points this at its C-ABI emulator)."""
    return bool(t.is_cuda)


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


_I, _P, _F, _D, _Z = c_int, c_void_p, c_float, c_double, c_size_t

class RowOp(ctypes.Structure):
    """struct ehgr_rowop (include/ehgr_b200.h)."""
    _fields_ = [("mode", ctypes.c_int32), ("relu6", ctypes.c_int32), ("in1", c_void_p), ("in2", c_void_p),
                ("scale", c_void_p), ("shift", c_void_p), ("ca", c_void_p), ("cb", c_void_p), ("cc", c_void_p),
                ("n_segment", ctypes.c_int32), ("fold", ctypes.c_int32), ("hw", ctypes.c_int32),
                ("shift_dir", ctypes.c_int32), ("cv_h", ctypes.c_int32), ("cv_w", ctypes.c_int32),
                ("cv_cin", ctypes.c_int32), ("cv_up", ctypes.c_int32)]


class ActionArgs(ctypes.Structure):
    """struct ehgr_action (include/ehgr_b200.h)."""
    _fields_ = ([(k, ctypes.c_int32) for k in ("n", "t", "h", "w", "c", "cr")] +
                [(k, c_void_p) for k in (
                    "shift_w", "p1_w", "p2_squeeze", "p2_conv1", "p2_expand", "p3_squeeze", "p3_conv1", "p3_expand",
                    "bn3_scale", "bn3_shift",
                    "mrow", "pool", "q", "qstats", "g1", "g2", "g3", "s", "u", "pi",
                    "dg1", "dgc", "dm", "dpool", "dd", "bn3_sums", "bn3_ca", "bn3_cb", "bn3_cc",
                    "d_shift_w", "d_p1_w", "d_p2_squeeze", "d_p2_conv1", "d_p2_expand", "d_p3_squeeze", "d_p3_conv1",
                    "d_p3_expand")])


class BnFin(ctypes.Structure):
    """struct ehgr_bnfin (include/ehgr_b200.h): BatchNorm finalisation run by the last CTA of the producing kernel."""
    _fields_ = [("gamma", c_void_p), ("beta", c_void_p), ("running_mean", c_void_p), ("running_var", c_void_p),
                ("scale", c_void_p), ("shift", c_void_p), ("mean", c_void_p), ("invstd", c_void_p), ("counter", c_void_p),
                ("count", c_longlong), ("momentum", c_float), ("eps", c_float), ("training", ctypes.c_int32)]


class BnBwd(ctypes.Structure):
    """struct ehgr_bnbwd: BatchNorm-backward coefficients computed by the last CTA of the reduction."""
    _fields_ = [("gamma", c_void_p), ("mean", c_void_p), ("invstd", c_void_p), ("ca", c_void_p), ("cb", c_void_p),
                ("cc", c_void_p), ("dgamma", c_void_p), ("dbeta", c_void_p), ("counter", c_void_p), ("count", c_longlong),
                ("training", ctypes.c_int32)]


_L = c_longlong
_R = ctypes.POINTER(RowOp)
_BF = ctypes.POINTER(BnFin)
_BB = ctypes.POINTER(BnBwd)
_A = ctypes.POINTER(ActionArgs)

# name -> argtypes for every int-returning symbol of include/ehgr_b200.h
# (tests/test_abi.py cross-checks this table against the header and the built library).
SIGNATURES = {
    "ehgr_temporal_shift_fwd": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "ehgr_temporal_shift_bwd": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "ehgr_pw_gemm": [_R, _P, _I, _P, _P, _P, _L, _I, _I, _I, _I, _P],
    "ehgr_pw_gemm_w16": [_R, _P, _P, _I, _P, _P, _P, _L, _I, _I, _I, _I, _P],
    "ehgr_pw_gemm_bn": [_R, _P, _P, _I, _P, _P, _P, _L, _I, _I, _I, _I, _BF, _P],
    "ehgr_pw_wgrad": [_R, _R, _P, _L, _I, _I, _I, _I, _P],
    "ehgr_conv3_pack": [_P, _P, _P, _I, _I, _I, _P],
    "ehgr_conv3_unpack_grad": [_P, _P, _I, _I, _P],
    "ehgr_upsample2_bwd": [_P, _P, _L, _I, _I, _I, _I, _P],
    "ehgr_upsample2_fwd": [_P, _P, _L, _I, _I, _I, _I, _P],
    "ehgr_depth_head_fwd": [_R, _P, _P, _P, _L, _I, _I, _P],
    "ehgr_depth_head_bwd": [_R, _P, _P, _P, _P, _P, _P, _L, _I, _I, _P],
    "ehgr_dw_fwd": [_R, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ehgr_dw_fwd_bn": [_R, _P, _P, _P, _I, _I, _I, _I, _I, _I, _BF, _P],
    "ehgr_dw_dgrad": [_R, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ehgr_dw_wgrad": [_R, _R, _P, _I, _I, _I, _I, _I, _I, _P],
    "ehgr_dw_bwd": [_R, _R, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ehgr_stem_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ehgr_stem_fwd_bn": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _BF, _P],
    "ehgr_stem_wgrad": [_R, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ehgr_bn_finalize": [_P, _L, _P, _P, _P, _P, _F, _F, _I, _P, _P, _P, _P, _I, _P],
    "ehgr_bn_bwd_reduce": [_P, _P, _P, _P, _I, _P, _L, _I, _I, _P],
    "ehgr_bn_bwd_finalize": [_P, _L, _P, _P, _P, _I, _P, _P, _P, _P, _P, _I, _P],
    "ehgr_bn_bwd_reduce_fin": [_P, _P, _P, _P, _I, _P, _L, _I, _I, _BB, _P],
    "ehgr_row_apply": [_R, _P, _P, _L, _I, _I, _P],
    "ehgr_normalize_u8": [_P, _P, _L, _I, _L, _P, _P, _F, _I, _P],
    "ehgr_sgd_step": [_P, _P, _P, _P, _P, _P, _I, _P, _F, _F, _L, _P, _D, _P, _P],
    "ehgr_ema_update": [_P, _P, _L, _D, _I, _P],
    "ehgr_temporal_pool_fwd": [_P, _P, _L, _I, _L, _I, _P],
    "ehgr_temporal_pool_bwd": [_P, _P, _P, _L, _I, _L, _I, _P],
    "ehgr_pool_fwd": [_R, _P, _I, _I, _I, _I, _P],
    "ehgr_pool_bwd": [_P, _P, _I, _I, _I, _I, _P],
    "ehgr_fc_consensus_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "ehgr_fc_consensus_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "ehgr_action_xs": [_A, _P, _P, _I, _P],
    "ehgr_action_gates": [_A, _P],
    "ehgr_action_bwd_reduce": [_A, _P, _P, _I, _P],
    "ehgr_action_bwd_small": [_A, _P],
    "ehgr_action_bwd_dxs": [_A, _P, _P, _P, _I, _P],
    "ehgr_action_fir_bwd": [_A, _P, _P, _P, _P, _I, _P],
    "ehgr_mtmm_loss": [_P, _P, _P, _P, _F, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ehgr_sd_loss": [_P, _P, _P, _F, _F, _F, _P, _P, _P, _I, _I, _L, _I, _P],
    # N3: torchvision ResNet bottleneck path (csrc/resnet.cu)
    "ehgr_stem7_fwd": [_P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _P],
    "ehgr_stem7_wgrad": [_R, _P, _P, _L, _I, _I, _I, _I, _I, _P],
    "ehgr_stem7_im2col": [_P, _P, _L, _I, _I, _I, _I, _I, _P],
    "ehgr_stem7_pack": [_P, _P, _I, _I, _I, _P],
    "ehgr_stem7_unpack_grad": [_P, _P, _I, _I, _P],
    "ehgr_maxpool3_fwd": [_R, _P, _P, _L, _I, _I, _I, _I, _P],
    "ehgr_maxpool3_bwd": [_P, _P, _P, _L, _I, _I, _I, _I, _P],
    "ehgr_subsample2_fwd": [_P, _P, _P, _L, _I, _I, _I, _I, _P],
    "ehgr_subsample2_bwd": [_P, _P, _L, _I, _I, _I, _I, _P],
    "ehgr_bn_add_relu": [_P, _P, _P, _P, _P, _L, _I, _I, _P],
    "ehgr_relu_bwd": [_P, _P, _P, _L, _I, _P],
}


def _declare(lib: ctypes.CDLL) -> None:
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int


# ------------------------------------------------------------------------------------------------
# launch helper + optional per-kernel CUDA-event timing (bench.py's roofline object)
# ------------------------------------------------------------------------------------------------
class KernelTimer:
    """When active, every `call()` is bracketed by CUDA events on the launching stream.  Event pairs
    are only read after the timed region (no synchronisation inside it)."""
    active = None  # dict name -> {"events": [(e0, e1)], "bytes": int, "flops": int}

    @classmethod
    def begin(cls):
        cls.active = {}
        return cls.active

    @classmethod
    def end(cls, rec):
        cls.active = None
        torch.cuda.synchronize()
        out = {}
        for name, r in rec.items():
            ms = sum(a.elapsed_time(b) for a, b in r["events"])
            out[name] = {"launches": len(r["events"]), "ms": ms, "bytes": r["bytes"], "flops": r["flops"]}
        return out


def call(name: str, *args, algo_bytes: int = 0, algo_flops: int = 0, tag: str = "") -> None:
    """Invoke C-ABI entry point `name` (args already raw pointers / ints) and raise on failure.  `tag` only labels the
    launch in the per-kernel timing table (e.g. the implicit-GEMM convolutions that share ehgr_pw_gemm_bn with the
    pointwise layers but are tensor-bound, not HBM-bound)."""
    fn = getattr(_load(), name)
    rec = KernelTimer.active
    if rec is None:
        check(fn(*args), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st = fn(*args)
    e1.record()
    check(st, name)
    r = rec.setdefault(name + tag, {"events": [], "bytes": 0, "flops": 0})
    r["events"].append((e0, e1))
    r["bytes"] += int(algo_bytes)
    r["flops"] += int(algo_flops)


def roofline_entry(kernel_times: dict, peaks: dict, peaks_kind: str, frames: int, traffic: dict = None):
    """bench.py `roofline` object for the dominant (largest total time) kernel of this library."""
    if not kernel_times:
        return None
    name, r = max(kernel_times.items(), key=lambda kv: kv[1]["ms"])
    total_ms = sum(v["ms"] for v in kernel_times.values())
    secs = r["ms"] / 1e3
    hbm = r["bytes"] / secs / 1e9 if secs > 0 else 0.0
    tfl = r["flops"] / secs / 1e12 if secs > 0 else 0.0
    ridge = peaks["bf16_tflops_sustained"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
    tensor_bound = r["bytes"] > 0 and (r["flops"] / r["bytes"]) > ridge
    entry = {
        "kernel": name, "bound": "tensor" if tensor_bound else "hbm",
        "achieved": round(tfl if tensor_bound else hbm, 2),
        "peak": peaks["bf16_tflops_sustained"] if tensor_bound else peaks["hbm_gbs"],
        "unit": "TFLOP/s" if tensor_bound else "GB/s",
        "traffic": (traffic or {}).get(name), "peak_source": peaks_kind,
        "launches": r["launches"], "avg_us": round(r["ms"] * 1e3 / max(1, r["launches"]), 2),
        "share_of_library_kernel_time": round(r["ms"] / total_ms, 4) if total_ms > 0 else None,
        "algo_bytes_per_launch": r["bytes"] // max(1, r["launches"]),
        "per_kernel": {k: {"launches": v["launches"], "ms": round(v["ms"], 3),
                           "GBps": round(v["bytes"] / (v["ms"] / 1e3) / 1e9, 1) if v["ms"] > 0 else None,
                           "TFLOPs": round(v["flops"] / (v["ms"] / 1e3) / 1e12, 2) if v["ms"] > 0 and v["flops"] else None}
                       for k, v in sorted(kernel_times.items(), key=lambda kv: -kv[1]["ms"])},
    }
    entry["frac"] = round(entry["achieved"] / entry["peak"], 4) if entry["peak"] else None
    return entry
