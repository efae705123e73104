"""The data-parallel training step on CPU, world_size 2 over gloo, with the kernels replaced by the C-ABI emulator
(tests/abi_emulator.py): MTMMTrainStep = forward, fused loss, backward with the gradient-sink protocol (the chain / the
ResNet function write parameter gradients straight into GradBuckets' flat buffer and report finished parameters, buckets
are all-reduced while backward continues), SGD.  Checked per rank: the averaged gradients equal the mean of the oracle's
gradients on the two ranks' shards, replicas stay in sync, and the parameters move by exactly -lr * (that gradient) on
the first step.  Cases: the MTMM step on TSM-MobileNetV2 (headline path) and on TSM-ResNet-50 (N3), and the SD step on
TSM-MobileNetV2 (its first bucket mixes several autograd Functions — the exit heads and the end of the backbone — which is
where a double-counted parameter let replicas diverge on 2 GPUs in round 2)."""
import contextlib
import io
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


CASES = ("mtmm-mobilenetv2", "mtmm-resnet50", "sd-mobilenetv2", "mtmmsd-mobilenetv2")


def _worker(rank, world, port, out):
    try:
        here = os.path.dirname(os.path.abspath(__file__))
        for p in (here, os.path.dirname(here)):
            if p not in sys.path:
                sys.path.insert(0, p)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.set_num_threads(4)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import abi_emulator
        import ehgr_b200 as E
        from conftest import check_grads_up_to_relu_flips
        from oracle import ref_oracle as O
        E._lib.call = abi_emulator.call
        E._lib.require_cuda = lambda *t: None
        E._lib.stream_ptr = lambda device=None: 0
        E._lib.on_gpu = lambda t: True         # the wrappers take the library's path for host tensors
        T, cls, size, lr = 2, 5, 64, 0.01
        for case in CASES:                     # one pair of processes runs every case (spawn + import cost paid once)
            try:
                _run_case(E, O, check_grads_up_to_relu_flips, case, rank, world, T, cls, size, lr)
            except Exception as e:
                raise RuntimeError(f"case {case}: {e!r}") from e
        if rank == 0:
            out.put("ok")
        dist.destroy_process_group()
    except Exception as e:      # surface the failure in the parent
        import traceback
        out.put("rank %d: %s\n%s" % (rank, e, traceback.format_exc()))
        raise


def _run_case(E, O, check_grads_up_to_relu_flips, case, rank, world, T, cls, size, lr):
    if True:                                   # (indentation kept from the single-case form)
        workload, backbone = case.split("-")
        resnet = backbone == "resnet50"
        sd_mode = workload == "sd"
        comb = workload == "mtmmsd"
        if comb:
            sd0 = O.build_mtmm_sd_state(cls, "tsm", 8, seed=4)
        elif sd_mode:
            sd0 = O.build_sd_state(cls, "tsm", 8, seed=4)
        else:
            sd0 = (O.build_resnet_mtmm_state(cls, "tsm", seed=4) if resnet else O.build_mtmm_state(cls, "tsm", 8, seed=4))
        common = dict(base_model=backbone, pretrain=None, dropout=0.5, partial_bn=False, is_shift=True, shift_div=8,
                      consensus_type='avg', fc_lr5=True, img_feature_dim=224, temporal_module='tsm', print_spec=False)
        with contextlib.redirect_stdout(io.StringIO()):
            model = (E.tsn_mtmm_sd.TSN(cls, T, 'RGB', modal='rgb_depth', **common) if comb else
                     E.tsn_sd.TSN(cls, T, 'RGB', **common) if sd_mode else
                     E.tsn_mtmm.TSN(cls, T, 'RGB', modal='rgb_depth', **common))
        model.load_state_dict(sd0, strict=True)
        model.train()
        for d in model.modules():
            if isinstance(d, torch.nn.Dropout):
                d.eval()
        step_cls = (E.train_step.MTMMSDTrainStep if comb else E.train_step.SDTrainStep if sd_mode else E.train_step.MTMMTrainStep)
        step = step_cls(model, lr=lr, momentum=0.9, weight_decay=0.0, compute_dtype=torch.float32, n_buckets=3)
        assert step.buckets.world == world
        before = {k: p.detach().clone() for k, p in model.named_parameters()}
        shards = [O.synthetic_clip_batch(2, T, size, cls, seed=20 + r) for r in range(world)]
        rgb, depth, labels = shards[rank]
        loss = step.run(rgb, labels) if sd_mode else step.run(rgb, depth, labels)
        assert torch.isfinite(loss)
        # reference: the mean over the ranks of the oracle's gradients on each shard (fp64)
        mean_grad = None
        for r in range(world):
            sd64 = O.clone_state(sd0, dtype=torch.float64)
            x5, dep, lab = shards[r][0].double(), shards[r][1].double(), shards[r][2]
            if comb:
                # train_mtmm_sd.py:240-293 at this resolution; beta scaled by the world size as in the SD step
                oo = O.mtmm_sd_forward(x5, sd64, T, "tsm", 8, True)
                gt = F.interpolate(dep.view(-1, 1, size, size), (size // 4, size // 4), mode='bilinear')
                (O.sd_loss(oo[:4], oo[4:8], lab, 0.1, 1e-6 * world, 3.0)[0] + 0.9 * 0.01 * F.mse_loss(oo[9], gt)).backward()
                g = {k: v.grad for k, v in sd64.items() if v.is_floating_point() and v.grad is not None}
                mean_grad = g if mean_grad is None else {k: mean_grad[k] + g[k] for k in g}
                continue
            if sd_mode:
                # SDTrainStep scales beta by the world size: the feature term is a SUM over the local batch (train_sd.py:191-193)
                oo = O.sd_forward(x5, sd64, T, "tsm", 8, True)
                O.sd_loss(oo[:4], oo[4:], lab, 0.1, 1e-6 * world, 3.0)[0].backward()
                g = {k: v.grad for k, v in sd64.items() if v.is_floating_point() and v.grad is not None}
                mean_grad = g if mean_grad is None else {k: mean_grad[k] + g[k] for k in g}
                continue
            if resnet:
                f = O.resnet_features(x5.view((-1, 3, size, size)), sd64, O.RESNET50_LAYERS, "tsm", T, 8, True)[-1]
                z = F.linear(f.mean((2, 3)), sd64["new_fc.weight"], sd64["new_fc.bias"])
                ol, od = z.view(-1, T, cls).mean(1), O.global_decoder(f, sd64, True)
            else:
                ol, od = O.mtmm_forward(x5, sd64, T, "tsm", 8, True)
            # train_mtmm.py:223-231 at this test's resolution (the oracle's mtmm_loss is written for 224 -> 56)
            gt = F.interpolate(dep.view(-1, 1, size, size), (size // 4, size // 4), mode='bilinear')
            (F.cross_entropy(ol, lab) + 0.01 * F.mse_loss(od, gt)).backward()
            g = {k: v.grad for k, v in sd64.items() if v.is_floating_point() and v.grad is not None}
            mean_grad = g if mean_grad is None else {k: mean_grad[k] + g[k] for k in g}
        mean_grad = {k: v / world for k, v in mean_grad.items()}
        # parameters outside the loss (local_decoder of the combined stage: its output is returned, not trained on) keep a
        # zero gradient and are left where they are
        used = [(k, p) for k, p in model.named_parameters() if k in mean_grad]
        assert all(float(p.grad.abs().sum()) == 0.0 for k, p in model.named_parameters() if k not in mean_grad)
        check_grads_up_to_relu_flips(used, mean_grad)
        # first SGD step (zero momentum buffer, no weight decay): p' = p - lr * lr_mult * averaged gradient
        mult = {id(p): g['lr_mult'] for g in model.get_optim_policies() for p in g['params']}
        for k, p in model.named_parameters():
            want = before[k] - lr * mult[id(p)] * p.grad
            assert torch.allclose(p.detach(), want, rtol=1e-5, atol=1e-7), k
        assert step.ranks_in_sync(), case


def test_data_parallel_steps_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
    msg = q.get(timeout=5)
    assert msg == "ok", msg
    assert all(p.exitcode == 0 for p in procs)
