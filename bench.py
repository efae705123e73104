#!/usr/bin/env python
"""bench.py — train clips/sec of the TSM-MobileNetV2 MTMM step (BASELINE.json metric), 8x224^2 clips.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--dtype bf16|fp32]

N>1 is launched by the driver as `python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N`
(one rank per GPU, NCCL).  A "step" is one pass of the hot path over one batch of synthetic clips:
zero-grad, forward (backbone + classifier + depth decoder), MTMM loss, backward, gradient all-reduce,
SGD update.  Workload at N=1: BASELINE.json configs[1] — MTMM stage-1, TSM-MobileNetV2 RGB +
pseudo-depth, batch 32/GPU, bf16 activations, 83 classes.

Prints ONE JSON line (rank 0).  `value` = whole-job clips/s with inputs resident in HBM; `e2e` = the
same through the public call with pinned HOST buffers (H2D of every batch and a D2H loss read inside
the timed region, next batch's copy overlapped on a side stream).
`--impl reference`: the reference's own algorithm on the host cores (the torch-fp32 oracle port of
the reference modules: the reference is Python and is not present on the GPU box), rank 0 only.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import sys
import threading
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "train clips/sec TSM-MBv2 8x224^2"
WORKLOAD_NAMES = {"mtmm": "MTMM stage-1 step (fwd+loss+bwd+allreduce+SGD), RGB+pseudo-depth",
                  "sd": "SD stage-2 step (fwd with 3 exit heads + SD loss + bwd + allreduce + SGD), RGB",
                  "mtmm_sd": "MTMM+SD combined step (3 exit heads + depth decoders + combined loss), RGB+pseudo-depth"}
UNIT = "clips/s"
T_SEG, SIZE, NUM_CLASS = 8, 224, 83


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="clips per GPU")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--temporal", default="tsm", choices=["tsm", "action", "none"])
    ap.add_argument("--workload", default="mtmm", choices=["mtmm", "sd", "mtmm_sd"],
                    help="mtmm: BASELINE configs[1] (the headline); sd: the stage-2 self-distillation step (configs[2]); "
                         "mtmm_sd: the combined stage of train_mtmm_sd.py")
    ap.add_argument("--classes", type=int, default=83, help="83 = EgoGesture (configs[0-2]), 25 = NvGesture (configs[3])")
    ap.add_argument("--cpu-clips", type=int, default=2, help="clips per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-input", choices=["uint8", "float32"], default="uint8",
                    help="host batch format of the end-to-end leg: uint8 frames normalised on the device, or fp32")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of one CUDA graph per step")
    ap.add_argument("--backbone", default="mobilenetv2", choices=["mobilenetv2", "resnet50"],
                    help="resnet50 (+ --temporal tsm|none, --segments 16): BASELINE configs[4], the N3 row — not the headline")
    ap.add_argument("--segments", type=int, default=T_SEG, help="frames per clip (8 for the headline configs, 16 for configs[4])")
    return ap.parse_args()


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period=0.1):
        super().__init__(daemon=True)
        self.period, self.samples, self.reasons, self.max_mhz = period, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def ncu_traffic():
    """DRAM bytes per launch of each library kernel, from the committed ncu capture of this same workload
    (profiles/r1_traffic.json; produced by the command named inside it) — never measured under the timer."""
    here = os.path.dirname(os.path.abspath(__file__))
    p = os.path.join(here, "profiles", "r2_traffic.json")
    if not os.path.exists(p):
        p = os.path.join(here, "profiles", "r1_traffic.json")
    try:
        with open(p) as f:
            return json.load(f)["bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        return None


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_step_rate(temporal: str, clips: int, steps: int, warmup: int, workload: str = "mtmm", classes: int = 83,
                            backbone: str = "mobilenetv2", segments: int = T_SEG):
    """clips/s of the reference step (fwd + loss + bwd + SGD) of `workload` on all host cores, fp32."""
    from oracle import ref_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    resnet = backbone == "resnet50"
    if resnet and workload != "mtmm":
        raise SystemExit("the CPU arm of --backbone resnet50 restates the MTMM step only")
    build = {"mtmm": O.build_mtmm_state, "sd": O.build_sd_state, "mtmm_sd": O.build_mtmm_sd_state}[workload]
    sd = O.clone_state(O.build_resnet_mtmm_state(classes, temporal, seed=0) if resnet else build(classes, temporal, 8, seed=0))
    params = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.SGD(params, lr=0.00125, momentum=0.9, weight_decay=5e-4)
    T_SEG = segments
    rgb, depth, labels = O.synthetic_clip_batch(clips, T_SEG, SIZE, classes, seed=0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if resnet:
            O.resnet_mtmm_train_step(sd, rgb, depth, labels, T_SEG, temporal, 8, True)
        elif workload == "mtmm":
            O.mtmm_train_step(sd, rgb, depth, labels, T_SEG, temporal, 8, True)
        elif workload == "sd":
            O.sd_train_step(sd, rgb, labels, T_SEG, temporal, 8, True)
        else:
            O.mtmm_sd_train_step(sd, rgb, depth, labels, T_SEG, temporal, 8, True)
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return clips / dt, dt, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the same K timed steps and W warm-up steps as the GPU arm (a CPU step of `--cpu-clips` clips takes ~0.2 s:
    # K = 20, W = 5 is a few seconds); only absurd requests are bounded so that the arm always ends within minutes
    steps = max(1, min(args.steps, 200))
    warm = max(0, min(args.warmup, 50))
    resnet = args.backbone == "resnet50"
    v, dt, cores = cpu_reference_step_rate(args.temporal, args.cpu_clips, steps, warm, args.workload, args.classes,
                                           args.backbone, args.segments)
    line = {
        "impl": "reference", "metric": (METRIC if not resnet else f"train clips/sec TSM-ResNet50 {args.segments}x224^2"),
        "value": round(v, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{WORKLOAD_NAMES[args.workload]}, {args.temporal.upper()}-"
                               f"{'ResNet50' if resnet else 'MobileNetV2'}, "
                               f"{args.segments}x224^2, {args.classes} classes; CPU sample = {args.cpu_clips} clips/step"},
        "cpu_baseline": {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps} steps of {args.cpu_clips} clips (fwd+loss+bwd+SGD), torch fp32, "
                                   f"{cores} threads"},
        "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import ehgr_b200
    from ehgr_b200 import _lib

    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    torch.manual_seed(1)  # train_mtmm.py:43 default seed; identical initial weights on every rank
    sd_mode = args.workload == "sd"
    NUM_CLASS = args.classes
    T_SEG = args.segments
    resnet = args.backbone == "resnet50"
    if resnet and (args.temporal == "action" or args.workload == "mtmm_sd"):
        raise SystemExit("--backbone resnet50 runs --temporal tsm|none with --workload mtmm|sd (ACTION covers widths <= 256 channels)")
    common = dict(is_shift=(args.temporal != "none"), partial_bn=False, base_model=args.backbone, shift_div=8, dropout=0.5,
                  img_feature_dim=224, pretrain=None, consensus_type='avg', fc_lr5=True,
                  temporal_module=("tsm" if args.temporal == "tsm" else "action"))
    with quiet():
        if sd_mode:
            model = ehgr_b200.tsn_sd.TSN(NUM_CLASS, T_SEG, 'RGB', **common)
        elif args.workload == "mtmm_sd":
            model = ehgr_b200.tsn_mtmm_sd.TSN(NUM_CLASS, T_SEG, 'RGB', modal='rgb_depth', **common)
        else:
            model = ehgr_b200.tsn_mtmm.TSN(NUM_CLASS, T_SEG, 'RGB', modal='rgb_depth', **common)
    model = model.to(dev)
    model.train()
    step_cls = {"sd": ehgr_b200.train_step.SDTrainStep, "mtmm_sd": ehgr_b200.train_step.MTMMSDTrainStep,
                "mtmm": ehgr_b200.train_step.MTMMTrainStep}[args.workload]
    step = step_cls(model, compute_dtype=dtype, use_graph=not args.no_graph)

    def pick(batch):          # the SD step takes (rgb, labels); the MTMM step (rgb, depth, labels)
        return (batch[0], batch[2]) if sd_mode else batch

    B = args.batch
    g = torch.Generator().manual_seed(100 + rank)
    n_host = 2
    host = [(torch.randn(B, T_SEG, 3, SIZE, SIZE, generator=g).pin_memory(),
             torch.rand(B, T_SEG, 1, SIZE, SIZE, generator=g).pin_memory(),
             torch.randint(0, NUM_CLASS, (B,), generator=g).pin_memory()) for _ in range(n_host)]
    resident = [tuple(t.to(dev) for t in h) for h in host]
    # end-to-end leg: the host batch is what a video loader holds before ToTorchFormatTensor / GroupNormalize —
    # uint8 frames (RGB) and uint8 pseudo-depth maps; the normalisation runs on the device (ehgr_normalize_u8).
    # --e2e-input float32 ships the already-normalised fp32 tensors instead (4x the bytes).
    if args.e2e_input == "uint8":
        host_e2e = [(torch.randint(0, 256, (B, T_SEG, 3, SIZE, SIZE), dtype=torch.uint8, generator=g).pin_memory(),
                     torch.randint(0, 256, (B, T_SEG, 1, SIZE, SIZE), dtype=torch.uint8, generator=g).pin_memory(),
                     h[2]) for h in host]
    else:
        host_e2e = host
    h2d_bytes = sum(t.numel() * t.element_size() for t in pick(host_e2e[0]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = {}

    def timed(fn, steps, tag=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        fn(steps)
        e1.record()
        if tag:
            host_ms[tag] = (time.perf_counter() - t0) * 1e3 / steps    # host time to ENQUEUE a step
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---- leg 1: inputs resident in HBM ----
    last_loss = {}

    def leg_resident(k):
        for i in range(k):
            last_loss["v"] = step.run(*pick(resident[i % n_host]))

    leg_resident(args.warmup)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _lib.launch_count()
    ms = timed(leg_resident, args.steps, "resident")
    launches = _lib.launch_count() - l0
    if step.use_graph:      # replayed launches never pass through the host entry points: count what was captured
        launches = step.launches_per_step * args.steps
    clocks = sampler.stop()
    # the same K steps once more with a CUDA-event pair around every library kernel (per-kernel durations
    # for the roofline object; kept out of the timed region above because 2 events per launch perturb it)
    graph_mode, step.use_graph = step.use_graph, False       # per-kernel events need eager launches
    prof = _lib.KernelTimer.begin()
    ms_prof = timed(leg_resident, args.steps)
    kernel_times = _lib.KernelTimer.end(prof)
    step.use_graph = graph_mode

    # ---- leg 2: end to end from pinned host memory, next batch prefetched on a copy stream ----
    copy_stream = torch.cuda.Stream(device=dev)

    def leg_e2e(k):
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(copy_stream):
            nxt = step.stage(*pick(host_e2e[0]))
        last = None
        for i in range(k):
            cur.wait_stream(copy_stream)
            batch = nxt
            for t in batch:
                t.record_stream(cur)
            if i + 1 < k:
                with torch.cuda.stream(copy_stream):
                    nxt = step.stage(*pick(host_e2e[(i + 1) % n_host]))
            loss = step.run(*batch)
            if last is not None:
                last.item()          # D2H read of the previous step's loss (keeps one step in flight)
            last = loss
        last.item()

    leg_e2e(max(args.warmup, 5))     # new input format: eager warm-up calls + graph capture happen here, untimed
    ms_e2e = timed(leg_e2e, args.steps)

    # replicas must hold identical parameters after all those steps with NCCL inside the captured graph
    in_sync = step.ranks_in_sync()
    if not in_sync:
        raise RuntimeError("data-parallel replicas diverged: parameter checksums differ between ranks")
    loss_value = float(last_loss["v"].item())
    if not (loss_value == loss_value and abs(loss_value) < 1e6):        # NaN / inf: the number would be meaningless
        raise RuntimeError(f"training loss is not finite ({loss_value}); refusing to report a throughput")
    if rank == 0:
        pk, pk_kind = peaks()
        clips = B * world * args.steps
        value = clips / (ms / 1e3)
        e2e = clips / (ms_e2e / 1e3)
        line = {
            "metric": (METRIC if not resnet else f"train clips/sec TSM-ResNet50 {T_SEG}x224^2"), "value": round(value, 2),
            "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"{WORKLOAD_NAMES[args.workload]}, {args.temporal.upper()}-"
                                   f"{'ResNet50' if resnet else 'MobileNetV2'}, {T_SEG}x224^2, "
                                   f"{NUM_CLASS} classes, train-mode BN",
                       "clips_per_gpu": B, "global_clips": B * world, "parallelism": f"dp{world}",
                       "launch": "one CUDA graph per step" if step.use_graph else "eager",
                       "l2": "activations >> 126 MB L2 (inputs larger than L2; no explicit flush)"},
            "e2e": {"value": round(e2e, 2), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "host_batch": ("uint8 frames + uint8 depth maps, normalised on the device" if args.e2e_input == "uint8"
                                   else "normalised fp32 tensors")},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": _lib.roofline_entry(kernel_times, pk, pk_kind, B * T_SEG, ncu_traffic() if not resnet else None),
            "roofline_traffic_source": ("static: DRAM bytes per launch from the committed ncu capture profiles/r2_traffic.json "
                                        "(same command, 1 GPU); not measured inside this run" if not resnet else
                                        "none: no ncu capture of the ResNet-50 shapes exists"),
            "peaks": pk_kind,
            "final_loss": round(loss_value, 5),
            "host_enqueue_ms_per_step": round(host_ms.get("resident", 0.0), 3),
            "ms_per_step_with_kernel_events": round(ms_prof / args.steps, 3),
        }
        if world > 1:
            line["ranks_in_sync"] = bool(in_sync)
        if not args.no_cpu_baseline and world == 1 and not (resnet and args.workload != "mtmm"):
            v, dt, cores = cpu_reference_step_rate(args.temporal, args.cpu_clips, 3, 1, args.workload, NUM_CLASS,
                                                   args.backbone, T_SEG)
            line["cpu_baseline"] = {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"3 steps of {args.cpu_clips} clips (fwd+loss+bwd+SGD), torch fp32 oracle "
                                              f"port of the reference modules, {cores} threads"}
        print(json.dumps(line), flush=True)
    if world > 1:
        # A captured graph holds NCCL kernels: drop it and drain the device before the communicator goes
        # away, then leave without the (occasionally hanging) communicator teardown.  The timer bounds
        # whatever part of this still blocks.
        sys.stdout.flush()
        wd = threading.Timer(30.0, lambda: os._exit(0))
        wd.daemon = True
        wd.start()
        step.invalidate_graph()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
