"""Fused chain (the reference's operator API on CUDA) against the oracle and the reference fixtures:
single InvertedResidual blocks, the whole TSN-MobileNetV2 forward+backward, the MTMM step."""
import contextlib
import io

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, rel_err
from oracle import ref_oracle as O

pytestmark = pytest.mark.gpu


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def _randomize(module, seed):
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = torch.rand(m.weight.shape, generator=g) * 0.5 + 0.75
            m.bias.data = torch.randn(m.bias.shape, generator=g) * 0.1
            m.running_mean.data = torch.randn(m.bias.shape, generator=g) * 0.1
            m.running_var.data = torch.rand(m.bias.shape, generator=g) * 0.5 + 0.75


@pytest.mark.parametrize("inp,oup,stride,t,temporal", [(24, 24, 1, 6, "none"), (24, 24, 1, 6, "tsm"), (32, 64, 2, 6, "none"),
                                                      (32, 16, 1, 1, "none"), (160, 160, 1, 6, "tsm"), (64, 96, 1, 6, "none"),
                                                      (32, 32, 1, 6, "action"), (96, 96, 1, 6, "action")])
@pytest.mark.parametrize("bn_train", [True, False])
def test_inverted_residual_block(inp, oup, stride, t, temporal, bn_train):
    import ehgr_b200 as E
    torch.manual_seed(0)
    with _quiet():
        blk = E.InvertedResidual(inp, oup, stride, t)
        if temporal == "tsm":
            blk.conv[0] = E.TemporalShift(blk.conv[0], n_segment=4, n_div=8)
        if temporal == "action":
            blk.conv[0] = E.Action(blk.conv[0], n_segment=4, shift_div=8)
            with torch.no_grad():
                g0 = torch.Generator().manual_seed(9)
                for k, p in blk.conv[0].named_parameters():
                    if k.startswith("action_") and "bn" not in k:
                        p.add_(torch.randn(p.shape, generator=g0) * 0.3)
    _randomize(blk, 3)
    sd0 = {"f." + k: v.clone() for k, v in blk.state_dict().items()}
    nt, hw = 8, 9
    x = torch.randn(nt, inp, hw, hw)
    g = torch.randn(nt, oup, (hw - 1) // stride + 1, (hw - 1) // stride + 1)
    # oracle, fp64
    sd = O.clone_state(sd0, dtype=torch.float64)
    x64 = x.double().requires_grad_(True)
    y64 = O.inverted_residual(x64, sd, "f", inp, oup, stride, t, temporal, 4, 8, bn_train)
    y64.backward(g.double())
    # CUDA
    blk = blk.cuda().train(bn_train)
    xd = x.cuda().requires_grad_(True)
    y = blk(xd)
    y.backward(g.cuda())
    assert rel_err(y, y64) < 1e-5
    assert rel_err(xd.grad, x64.grad) < 2e-5
    gmax = max(v.grad.abs().max().item() for v in sd.values() if v.grad is not None)
    for k, p in blk.named_parameters():
        ref = sd["f." + k].grad
        assert ((p.grad.cpu().double() - ref).abs().max().item() <= 2e-5 * max(ref.abs().max().item(), 1e-3 * gmax)), k
    if bn_train:
        new = blk.state_dict()
        for k in new:
            if "running" in k:
                assert rel_err(new[k], sd["f." + k]) < 1e-5, k
            if "num_batches" in k:
                assert int(new[k]) == 1


@pytest.mark.parametrize("name", ["c32", "c24", "c160"])
def test_action_module_against_reference_fixture(name):
    """Stand-alone Action (fused xs / gates / backward kernels + library `net`) against the outputs and
    gradients of the reference module (tests/golden/action.npz)."""
    import ehgr_b200 as E
    z = np.load(GOLDEN / "action.npz")
    c, h, T, n, train_bn = (int(v) for v in z[name + "_meta"])
    rs = np.random.RandomState(11 + c)
    sd = {}
    O.action_state(sd, "m", c, 8, rs)
    O._conv_entry(sd, "m.net.weight", (6 * c, c, 1, 1), rs)
    with _quiet():
        mod = E.Action(torch.nn.Conv2d(c, 6 * c, 1, bias=False), n_segment=T, shift_div=8)
    mod.load_state_dict({k[2:]: v for k, v in sd.items()}, strict=True)
    mod = mod.cuda().train(bool(train_bn))
    x = torch.from_numpy(z[name + "_x"]).cuda().requires_grad_(True)
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        y = mod(x)
        y.backward(torch.from_numpy(z[name + "_g"]).cuda())
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    assert rel_err(y, torch.from_numpy(z[name + "_y"])) < 1e-5
    assert rel_err(x.grad, torch.from_numpy(z[name + "_gx"])) < 2e-5
    gmax = max(np.abs(z[k]).max() for k in z.files if k.startswith(name + "_grad_"))
    for k in z.files:
        if k.startswith(name + "_grad_"):
            pk = k[len(name + "_grad_"):]
            got = dict(mod.named_parameters())[pk].grad.cpu().double()
            ref = torch.from_numpy(z[k]).double()
            assert (got - ref).abs().max().item() <= 3e-5 * max(ref.abs().max().item(), 1e-3 * gmax), pk
    if train_bn:
        assert rel_err(mod.action_p3_bn1.running_mean, torch.from_numpy(z[name + "_rm"])) < 1e-5
        assert rel_err(mod.action_p3_bn1.running_var, torch.from_numpy(z[name + "_rv"])) < 1e-5


def _digest(g):
    g = g.detach().double().flatten().cpu()
    idx = torch.linspace(0, g.numel() - 1, steps=min(16, g.numel())).long()
    return np.concatenate([[g.sum().item(), g.abs().sum().item()], g[idx].numpy()])


GRAD_NOISE_FACTOR = 4.0


def _tsn(temporal):
    import ehgr_b200 as E
    with _quiet():
        return E.TSN(83, 8, 'RGB', base_model='mobilenetv2', pretrain=None, dropout=0.5, partial_bn=False,
                     is_shift=(temporal != "none"), shift_div=8, consensus_type='avg', fc_lr5=True, img_feature_dim=224,
                     temporal_module=("tsm" if temporal == "tsm" else "action"))


@pytest.mark.parametrize("temporal", ["none", "tsm", "action"])
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_tsn_against_reference_fixture(temporal, mode):
    """Whole network (4 clips at 96x96), fp32 kernels, against the live-reference fixture; the arbiter is the reference's
    fp64 run.
      * logits, loss, running statistics: the flat 1e-5 of north_star, in both BatchNorm modes;
      * gradients with FROZEN BatchNorm (eval): flat 1e-5 as well;
      * gradients with batch-statistics BatchNorm (train): the reference's OWN fp32 run is 0.6-1.2e-2 away from its fp64
        run on this fixture (52 BatchNorm layers back-to-back with random weights: every layer's backward subtracts two
        nearly equal means) — no fp32 implementation can hold 1e-5 there, PyTorch's included.  The bound is
        GRAD_NOISE_FACTOR x that measured reference noise; the 1e-5 line for train-mode gradients is held per block
        (test_inverted_residual_block) where the problem is well conditioned."""
    z = np.load(GOLDEN / "tsn_mbv2.npz")
    tag = f"{temporal}_{mode}"
    m = _tsn(temporal)
    m.load_state_dict(O.build_tsn_state(83, temporal, 8, seed=5), strict=True)
    m = m.cuda().train(mode == "train")
    for d in m.modules():
        if isinstance(d, torch.nn.Dropout):
            d.eval()
    rgb, _, labels = O.synthetic_clip_batch(4, 8, 96, 83, seed=3)
    logits = m(rgb.cuda())
    ref64 = torch.from_numpy(z[tag + "_logits64"])
    assert rel_err(logits, ref64) < 1e-5
    loss = F.cross_entropy(logits, labels.cuda())
    assert abs(loss.item() - float(z[tag + "_loss64"])) < 1e-5
    loss.backward()
    scale = max(np.abs(z[k][2:]).max() for k in z.files if k.startswith(tag + "_g64_"))
    ref_noise = max(np.abs(z[k][2:] - z[k.replace("_g_", "_g64_")][2:]).max() for k in z.files if k.startswith(tag + "_g_")) / scale
    worst = 0.0
    params = dict(m.named_parameters())
    for k in z.files:
        if k.startswith(tag + "_g64_"):
            got = _digest(params[k[len(tag + "_g64_"):]].grad)
            worst = max(worst, np.abs(got[2:] - z[k][2:]).max() / scale)
    # full gradient tensors of the first and the last layer (stem weight, classifier bias, 8 classifier rows)
    for k in z.files:
        if k.startswith(tag + "_gfull64_"):
            name = k[len(tag + "_gfull64_"):]
            got = params[name].grad.detach().cpu().double()
            ref = torch.from_numpy(z[k]).double()
            worst = max(worst, float((got[:ref.shape[0]] - ref).abs().max()) / scale)
    print(f"[{tag}] gradient error {worst:.3e}, reference fp32-vs-fp64 {ref_noise:.3e}")
    assert worst < (1e-5 if mode == "eval" else GRAD_NOISE_FACTOR * ref_noise), (worst, ref_noise)
    if mode == "train":
        sd = m.state_dict()
        for k in z.files:
            if k.startswith(tag + "_rs64_"):
                assert rel_err(sd[k[len(tag + "_rs64_"):]], torch.from_numpy(z[k])) < 1e-5


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
def test_mtmm_step_against_oracle(dtype, tol):
    """Config #1 of BASELINE.json scaled to one clip: MTMM forward, loss, backward at 224x224."""
    import ehgr_b200 as E
    sd0 = O.build_mtmm_state(83, "tsm", 8, seed=2)
    with _quiet():
        model = E.tsn_mtmm.TSN(83, 8, 'RGB', is_shift=True, partial_bn=False, base_model='mobilenetv2', shift_div=8,
                               dropout=0.5, img_feature_dim=224, pretrain=None, consensus_type='avg', fc_lr5=True,
                               modal='rgb_depth', temporal_module='tsm')
    model.load_state_dict(sd0, strict=True)
    model = model.cuda().train()
    for d in model.modules():
        if isinstance(d, torch.nn.Dropout):
            d.eval()
    rgb, depth, labels = O.synthetic_clip_batch(1, 8, 224, 83, seed=4)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False      # (the oracle side, when it runs on the GPU; our step has no library convolution)
    try:
        with E.fused.compute_dtype(dtype):
            logits, dpred = model(rgb.cuda())
            loss, _ = E.losses.mtmm_loss(logits, labels.cuda(), dpred, depth.cuda())
        loss.backward()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    # arbiter: the oracle in fp64.  Yardstick: the SAME oracle (plain PyTorch ops) in the precision under
    # test — fp32 on the CPU, bf16 autocast on the GPU.  This random-weight, train-mode-BN, single-clip case
    # is ill-conditioned (PyTorch's own bf16 run is ~35% off in the last feature map), so
    # the claim is "no worse than PyTorch at the same precision", per-op tests hold the 1e-5 / 2e-2 lines.
    sd = O.clone_state(sd0, dtype=torch.float64)
    oloss, ologits, odpred = O.mtmm_train_step(sd, rgb.double(), depth.double(), labels, 8, "tsm", 8, True)
    if dtype == torch.float32:
        sdy = O.clone_state(sd0)
        yloss, ylogits, ydpred = O.mtmm_train_step(sdy, rgb, depth, labels, 8, "tsm", 8, True)
    else:
        sdy = {k: (v.detach().float().cuda().requires_grad_(v.requires_grad) if v.is_floating_point() else v.cuda())
               for k, v in O.clone_state(sd0).items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            yloss, ylogits, ydpred = O.mtmm_train_step(sdy, rgb.cuda(), depth.cuda(), labels.cuda(), 8, "tsm", 8, True)
    gmax = max(v.grad.abs().max().item() for v in sd.values() if v.grad is not None)

    def grad_err(named):
        return max(((g.detach().cpu().double() - sd[k].grad).abs().max().item() / gmax) for k, g in named)

    ours = {"logits": rel_err(logits, ologits), "depth": (dpred.detach().cpu().double() - odpred).abs().max().item(),
            "loss": abs(loss.item() - oloss.item()),
            "grad": grad_err((k, p.grad) for k, p in model.named_parameters())}
    ref = {"logits": rel_err(ylogits, ologits), "depth": (ydpred.detach().cpu().double() - odpred).abs().max().item(),
           "loss": abs(float(yloss) - oloss.item()),
           "grad": grad_err((k, v.grad) for k, v in sdy.items() if v.is_floating_point() and v.grad is not None)}
    for k in ours:
        assert ours[k] <= max(3.0 * ref[k], tol), (k, ours, ref)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
def test_sd_step_against_oracle(dtype, tol):
    """SD stage-2 step (train_sd.py:217-282) on TSM-MobileNetV2 with the three exit heads: fused chain
    with taps + fused SD loss kernel, against the oracle; yardstick = the oracle at the same precision."""
    import ehgr_b200 as E
    sd0 = O.build_sd_state(83, "tsm", 8, seed=6)
    with _quiet():
        model = E.tsn_sd.TSN(83, 8, 'RGB', is_shift=True, partial_bn=False, base_model='mobilenetv2', shift_div=8,
                             dropout=0.5, img_feature_dim=224, pretrain=None, consensus_type='avg', fc_lr5=True,
                             temporal_module='tsm')
    model.load_state_dict(sd0, strict=True)
    model = model.cuda().train()
    for d in model.modules():
        if isinstance(d, torch.nn.Dropout):
            d.eval()
    rgb, _, labels = O.synthetic_clip_batch(1, 8, 224, 83, seed=7)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False      # the exit heads still run on library convolutions
    try:
        with E.fused.compute_dtype(dtype):
            outs = model(rgb.cuda())
            total, terms = E.losses.sd_loss(outs[:4], outs[4:], labels.cuda(), 0.1, 1e-6, 3.0)
        total.backward()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert [tuple(o.shape) for o in outs] == [(1, 83)] * 4 + [(8, 1280, 1, 1)] * 4
    sd = O.clone_state(sd0, dtype=torch.float64)
    ototal, oterms = O.sd_train_step(sd, rgb.double(), labels)
    if dtype == torch.float32:
        sdy = O.clone_state(sd0)
        ytotal, _ = O.sd_train_step(sdy, rgb, labels)
    else:
        sdy = {k: (v.detach().float().cuda().requires_grad_(v.requires_grad) if v.is_floating_point() else v.cuda())
               for k, v in O.clone_state(sd0).items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ytotal, _ = O.sd_train_step(sdy, rgb.cuda(), labels.cuda())
    gmax = max(v.grad.abs().max().item() for v in sd.values() if v.grad is not None)

    def grad_err(named):
        return max(((g.detach().cpu().double() - sd[k].grad).abs().max().item() / gmax) for k, g in named)

    ours = {"loss": abs(total.item() - ototal.item()) / abs(ototal.item()),
            "grad": grad_err((k, p.grad) for k, p in model.named_parameters())}
    ref = {"loss": abs(float(ytotal) - ototal.item()) / abs(ototal.item()),
           "grad": grad_err((k, v.grad) for k, v in sdy.items() if v.is_floating_point() and v.grad is not None)}
    for k in ours:
        assert ours[k] <= max(3.0 * ref[k], tol), (k, ours, ref)


def test_train_step_cuda_graph_matches_eager():
    """MTMMTrainStep in CUDA-graph mode (two eager warm-up calls, capture, replays) walks the same trajectory
    as the eager step.  Atomics make every run order-dependent and this random-weight problem amplifies
    that, so the yardstick is a SECOND eager run: graph-vs-eager may not exceed eager-vs-eager by more than
    a small factor.  fp32 storage keeps the noise floor low."""
    import ehgr_b200 as E

    def make():
        sd0 = O.build_mtmm_state(83, "tsm", 8, seed=6)
        with _quiet():
            model = E.tsn_mtmm.TSN(83, 8, 'RGB', is_shift=True, partial_bn=False, base_model='mobilenetv2', shift_div=8,
                                   dropout=0.5, img_feature_dim=224, pretrain=None, consensus_type='avg', fc_lr5=True,
                                   modal='rgb_depth', temporal_module='tsm')
        model.load_state_dict(sd0, strict=True)
        model = model.cuda().train()
        for d in model.modules():
            if isinstance(d, torch.nn.Dropout):
                d.eval()
        return model

    # two eager warm-up calls, the capture + first replay, one more replay: short enough that the chaotic
    # amplification of atomics-order noise (1e-3 in the loss after five steps of this problem) stays small
    batches = [tuple(t.cuda() for t in O.synthetic_clip_batch(2, 8, 64, 83, seed=20 + i)) for i in range(4)]
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        runs = {}
        for mode in ("eager", "eager2", "graph"):
            model = make()
            step = E.train_step.MTMMTrainStep(model, lr=1e-4, compute_dtype=torch.float32, use_graph=(mode == "graph"))
            losses = [float(step.run(*b).item()) for b in batches]
            runs[mode] = (losses, {k: v.detach().clone() for k, v in model.state_dict().items()})
            if mode == "graph":
                assert step._graph is not None and step.launches_per_step > 100
    finally:
        torch.backends.cudnn.allow_tf32 = old

    def dist_to_eager(mode):
        la, lb = runs["eager"][0], runs[mode][0]
        dl = max(abs(a - b) / abs(a) for a, b in zip(la, lb))
        sa, sb = runs["eager"][1], runs[mode][1]
        for k in sa:
            if not sa[k].is_floating_point():
                assert torch.equal(sa[k], sb[k]), k
        dp = max(rel_err(sb[k], sa[k]) for k in sa if sa[k].is_floating_point())
        return dl, dp

    noise_l, noise_p = dist_to_eager("eager2")
    dl, dp = dist_to_eager("graph")
    # successive batches differ by several percent in loss, so a replay on stale inputs or with a missing
    # update would show up far above these bounds
    assert dl <= max(10 * noise_l, 3e-3), (dl, noise_l)
    assert dp <= max(10 * noise_p, 2e-2), (dp, noise_p)


def test_sd_train_step_runs_in_graph_mode():
    """SDTrainStep (three exit heads + fused SD loss) captured in a CUDA graph: finite, changing losses and
    moving parameters over replays with different batches."""
    import ehgr_b200 as E
    sd0 = O.build_sd_state(83, "tsm", 8, seed=3)
    with _quiet():
        model = E.tsn_sd.TSN(83, 8, 'RGB', is_shift=True, partial_bn=False, base_model='mobilenetv2', shift_div=8,
                             dropout=0.5, img_feature_dim=224, pretrain=None, consensus_type='avg', fc_lr5=True,
                             temporal_module='tsm')
    model.load_state_dict(sd0, strict=True)
    model = model.cuda().train()
    step = E.train_step.SDTrainStep(model, lr=1e-3, compute_dtype=torch.bfloat16, use_graph=True)
    w0 = model.base_model.features[1].conv[0].weight.detach().clone()
    losses = []
    for i in range(5):
        rgb, _depth, labels = O.synthetic_clip_batch(2, 8, 64, 83, seed=40 + i)
        losses.append(float(step.run(rgb.cuda(), labels.cuda()).item()))
    assert step._graph is not None and step.launches_per_step > 100
    assert all(l == l and abs(l) < 1e4 for l in losses), losses
    assert len({round(l, 4) for l in losses}) > 1
    assert not torch.equal(w0, model.base_model.features[1].conv[0].weight)


def test_flat_sgd_matches_torch_sgd():
    """FlatSGD (one kernel over flat buffers, device-side learning rate) against torch.optim.SGD with the same
    policy groups over several steps, including a learning-rate change."""
    import ehgr_b200 as E
    torch.manual_seed(0)

    def make():
        ps = [torch.nn.Parameter(torch.randn(s, generator=torch.Generator().manual_seed(i)).cuda())
              for i, s in enumerate([(7, 3), (5,), (2, 2, 2), (11,), (1,), (33, 9)])]
        pol = [{'params': [ps[0], ps[5]], 'lr_mult': 1, 'decay_mult': 1, 'name': 'w'},
               {'params': [ps[1], ps[3]], 'lr_mult': 2, 'decay_mult': 0, 'name': 'b'},
               {'params': [ps[2], ps[4]], 'lr_mult': 5, 'decay_mult': 1, 'name': 'fc'},
               {'params': [], 'lr_mult': 1, 'decay_mult': 1, 'name': 'empty'}]
        return ps, pol

    ps_a, pol_a = make()
    ps_b, pol_b = make()
    gb = E.train_step.GradBuckets(ps_a, n_buckets=2)
    opt_a = E.train_step.FlatSGD(pol_a, gb, lr=0.1, momentum=0.9, weight_decay=5e-4)
    groups = [dict(params=g['params'], lr=0.1 * g['lr_mult'], weight_decay=5e-4 * g['decay_mult']) for g in pol_b if g['params']]
    opt_b = torch.optim.SGD(groups, momentum=0.9)
    for step in range(5):
        if step == 3:
            E.train_step.adjust_learning_rate(0.1, opt_a, 1, [1])      # gamma 0.1 from epoch 1 on
            for g, pg in zip(opt_b.param_groups, [p for p in pol_b if p['params']]):
                g['lr'] = 0.1 * 0.1 * pg['lr_mult']
        gb.zero()
        for i, (pa, pb) in enumerate(zip(ps_a, ps_b)):
            gr = torch.randn(pa.shape, generator=torch.Generator().manual_seed(100 * step + i)).cuda()
            pa.grad.copy_(gr)
            pb.grad = gr.clone()
        opt_a.step()
        opt_b.step()
        for pa, pb in zip(ps_a, ps_b):
            assert rel_err(pa.detach(), pb.detach()) < 1e-6, step
    assert all(p.data_ptr() >= opt_a.flat_p.data_ptr() for p in ps_a)


def test_training_reduces_the_loss():
    """End-to-end sanity of the whole step (fused forward/backward, gradient sink, FlatSGD, CUDA graph): a few
    dozen steps on one small fixed batch must drive the MTMM loss down substantially."""
    import ehgr_b200 as E
    sd0 = O.build_mtmm_state(83, "tsm", 8, seed=9)
    with _quiet():
        model = E.tsn_mtmm.TSN(83, 8, 'RGB', is_shift=True, partial_bn=False, base_model='mobilenetv2', shift_div=8,
                               dropout=0.5, img_feature_dim=224, pretrain=None, consensus_type='avg', fc_lr5=True,
                               modal='rgb_depth', temporal_module='tsm')
    model.load_state_dict(sd0, strict=True)
    model = model.cuda().train()
    for d in model.modules():
        if isinstance(d, torch.nn.Dropout):
            d.eval()
    step = E.train_step.MTMMTrainStep(model, lr=2e-3, compute_dtype=torch.bfloat16, use_graph=True)
    batch = tuple(t.cuda() for t in O.synthetic_clip_batch(4, 8, 64, 83, seed=77))
    losses = [float(step.run(*batch).item()) for _ in range(40)]
    assert all(l == l for l in losses), losses
    first, last = sum(losses[:3]) / 3, sum(losses[-3:]) / 3
    assert last < 0.6 * first, (first, last, losses[::5])


def test_train_step_uint8_batch_equals_normalised_float_batch():
    """The uint8 entry of the train step (frames normalised on the device) computes the same loss as feeding the
    tensors the reference's CPU transforms would have produced."""
    import ehgr_b200 as E
    g = torch.Generator().manual_seed(3)
    rgb8 = torch.randint(0, 256, (2, 8, 3, 64, 64), dtype=torch.uint8, generator=g)
    dep8 = torch.randint(0, 256, (2, 8, 1, 64, 64), dtype=torch.uint8, generator=g)
    labels = torch.randint(0, 83, (2,), generator=g)
    rgbf = rgb8.float().div(255)
    for ci, (m, s) in enumerate(zip([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])):
        rgbf[:, :, ci].sub_(m).div_(s)
    depf = dep8.float().div(255)
    losses = []
    for batch in ((rgb8, dep8, labels), (rgbf, depf, labels)):
        sd0 = O.build_mtmm_state(83, "tsm", 8, seed=12)
        with _quiet():
            model = E.tsn_mtmm.TSN(83, 8, 'RGB', is_shift=True, partial_bn=False, base_model='mobilenetv2', shift_div=8,
                                   dropout=0.5, img_feature_dim=224, pretrain=None, consensus_type='avg', fc_lr5=True,
                                   modal='rgb_depth', temporal_module='tsm')
        model.load_state_dict(sd0, strict=True)
        model = model.cuda().train()
        for d in model.modules():
            if isinstance(d, torch.nn.Dropout):
                d.eval()
        step = E.train_step.MTMMTrainStep(model, lr=1e-4, compute_dtype=torch.float32)
        losses.append(float(step.run(*(t.cuda() for t in batch)).item()))
    assert abs(losses[0] - losses[1]) <= 1e-5 * abs(losses[1]), losses


def test_flat_ema_follows_the_reference_emawrapper_bit_exact():
    """MTMMTrainStep(ema_decay=...) keeps `model_ema` the way the reference's train loop does (EMAWrapper.update after
    every optimizer.step(), train_mtmm.py:242-245) with the parameter part folded into the SGD kernel.  The yardstick is
    the oracle's ema_update (pinned bit-exact to the reference class in tests/test_oracle_golden.py) applied to the
    model's state_dict after every step."""
    import ehgr_b200 as E
    sd0 = O.build_mtmm_state(83, "tsm", 8, seed=31)
    with _quiet():
        model = E.tsn_mtmm.TSN(83, 8, 'RGB', is_shift=True, partial_bn=False, base_model='mobilenetv2', shift_div=8,
                               dropout=0.5, img_feature_dim=224, pretrain=None, consensus_type='avg', fc_lr5=True,
                               modal='rgb_depth', temporal_module='tsm')
    model.load_state_dict(sd0, strict=True)
    model = model.cuda().train()
    step = E.train_step.MTMMTrainStep(model, lr=1e-2, compute_dtype=torch.float32, ema_decay=0.9)
    ema_ref = {k: v.detach().clone() for k, v in model.state_dict().items()}
    assert set(step.ema.state_dict().keys()) == set(ema_ref.keys())
    for i in range(3):
        batch = tuple(t.cuda() for t in O.synthetic_clip_batch(2, 8, 64, 83, seed=50 + i))
        step.run(*batch)
        O.ema_update(ema_ref, model.state_dict(), 0.9)
    got = step.ema.state_dict()
    moved = 0
    for k, v in got.items():
        assert v.dtype == ema_ref[k].dtype and torch.equal(v, ema_ref[k]), k
        moved += int(not torch.equal(v, sd0[k].to(v.device)))
    assert moved > 200                                  # parameters, running statistics and counters all moved
    # the EMA model is a working module: same forward entry as the reference wrapper
    with torch.no_grad():
        out = step.ema(batch[0])
    assert out[0].shape == (2, 83)


@pytest.mark.parametrize("name", ["sep_a", "sep_b", "sep_c"])
def test_sepconv_on_the_fused_chain_against_reference_fixture(name):
    """SepConv (models/models_SD.py:81-101) as four stages of the fused chain (bare depthwise convolutions, pointwise
    GEMMs, plain-ReLU row operands) against outputs, input gradient, every parameter gradient and the running
    statistics of the live reference module (tests/golden/heads.npz), fp32 storage, 1e-5 / 2e-5."""
    import ehgr_b200 as E
    z = np.load(GOLDEN / "heads.npz")
    ci, co, h, n, train_bn = (int(v) for v in z[name + "_meta"])
    sd = {}
    O.sepconv_state(sd, "m", ci, co, np.random.RandomState(501 + ci))
    mod = E.tsn_sd.SepConv(ci, co)
    mod.load_state_dict({k[2:]: v for k, v in sd.items()}, strict=True)
    mod = mod.cuda().train(bool(train_bn))
    x = torch.from_numpy(z[name + "_x"]).cuda().requires_grad_(True)
    l0 = E._lib.launch_count()
    y = mod(x)
    y.backward(torch.from_numpy(z[name + "_g"]).cuda())
    assert E._lib.launch_count() - l0 >= 12                  # our kernels ran (no library convolution path)
    assert rel_err(y, torch.from_numpy(z[name + "_y"])) < 1e-5
    assert rel_err(x.grad, torch.from_numpy(z[name + "_gx"])) < 2e-5
    gmax = max(np.abs(z[k]).max() for k in z.files if k.startswith(name + "_grad_"))
    for k in z.files:
        if k.startswith(name + "_grad_"):
            pk = k[len(name + "_grad_"):]
            got = dict(mod.named_parameters())[pk].grad.cpu().double()
            ref = torch.from_numpy(z[k]).double()
            assert (got - ref).abs().max().item() <= 3e-5 * max(ref.abs().max().item(), 1e-3 * gmax), pk
        if k.startswith(name + "_buf_") and train_bn:
            assert rel_err(dict(mod.named_buffers())[k[len(name + "_buf_"):]], torch.from_numpy(z[k])) < 1e-5, k


def test_mtmm_sd_loss_matches_reference_fixture():
    """Combined MTMM+SD loss (train_mtmm_sd.py:240-293; fixture = the reference statements with its own kd / feature
    functions, tests/golden/heads.npz ms_*): total, and the gradient of every input, within 1e-5."""
    import ehgr_b200 as E
    z = np.load(GOLDEN / "heads.npz")
    lg = [torch.from_numpy(z[f"ms_logits{i}"]).cuda().requires_grad_(True) for i in range(4)]
    ft = [torch.from_numpy(z[f"ms_feat{i}"]).cuda().requires_grad_(True) for i in range(4)]
    pred = torch.from_numpy(z["ms_gpred_in"]).cuda().requires_grad_(True)
    depth = torch.from_numpy(z["ms_depth"].astype(np.float32)).cuda()
    labels = torch.from_numpy(z["ms_labels"]).cuda()
    total, terms, mse = E.losses.mtmm_sd_loss(lg, ft, pred, depth, labels, 0.1, 1e-6, 3.0)
    assert abs(total.item() - float(z["ms_total"])) < 1e-5 * abs(float(z["ms_total"]))
    assert abs((terms[0] + 0.01 * mse).item() - float(z["ms_loss"])) < 1e-5      # `loss` of the reference = CE + 0.01 MSE
    total.backward()
    for i in range(4):
        assert rel_err(lg[i].grad, torch.from_numpy(z[f"ms_glogits{i}"])) < 1e-5
        if i:
            assert rel_err(ft[i].grad, torch.from_numpy(z[f"ms_gfeat{i}"])) < 1e-5
    assert ft[0].grad is None or float(ft[0].grad.abs().max()) == 0.0
    assert rel_err(pred.grad, torch.from_numpy(z["ms_gpred"])) < 1e-5


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
def test_mtmm_sd_step_against_oracle(dtype, tol):
    """The combined MTMM+SD wrapper on TSM-MobileNetV2 (models/models_MTMM_SD.py:431-532 generalised: ten outputs for
    modal='rgb_depth') + combined loss, one clip at 224x224, against the oracle (its decoders and loss are pinned to the
    live reference); yardstick = the oracle at the same precision, as for the MTMM and SD steps."""
    import ehgr_b200 as E
    sd0 = O.build_mtmm_sd_state(83, "tsm", 8, seed=8)
    with _quiet():
        model = E.tsn_mtmm_sd.TSN(83, 8, 'RGB', is_shift=True, partial_bn=False, base_model='mobilenetv2', shift_div=8,
                                  dropout=0.5, img_feature_dim=224, pretrain=None, consensus_type='avg', fc_lr5=True,
                                  modal='rgb_depth', temporal_module='tsm')
    model.load_state_dict(sd0, strict=True)
    model = model.cuda().train()
    for d in model.modules():
        if isinstance(d, torch.nn.Dropout):
            d.eval()
    rgb, depth, labels = O.synthetic_clip_batch(1, 8, 224, 83, seed=9)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False      # the ConvTranspose decoders run on library kernels
    try:
        with E.fused.compute_dtype(dtype):
            outs = model(rgb.cuda())
            total, terms, mse = E.losses.mtmm_sd_loss(outs[:4], outs[4:8], outs[9], depth.cuda(), labels.cuda())
        total.backward()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert [tuple(o.shape) for o in outs] == [(1, 83)] * 4 + [(8, 1280, 1, 1)] * 4 + [(8, 1, 224, 224), (8, 1, 56, 56)]
    sd = O.clone_state(sd0, dtype=torch.float64)
    ototal, _oloss, _ = O.mtmm_sd_train_step(sd, rgb.double(), depth.double(), labels)
    if dtype == torch.float32:
        sdy = O.clone_state(sd0)
        ytotal, _, _ = O.mtmm_sd_train_step(sdy, rgb, depth, labels)
    else:
        sdy = {k: (v.detach().float().cuda().requires_grad_(v.requires_grad) if v.is_floating_point() else v.cuda())
               for k, v in O.clone_state(sd0).items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ytotal, _, _ = O.mtmm_sd_train_step(sdy, rgb.cuda(), depth.cuda(), labels.cuda())
    gmax = max(v.grad.abs().max().item() for v in sd.values() if v.grad is not None)

    def grad_err(named):
        return max(((g.detach().cpu().double() - sd[k].grad).abs().max().item() / gmax) for k, g in named if g is not None and sd[k].grad is not None)

    ours = {"loss": abs(total.item() - ototal.item()) / abs(ototal.item()),
            "grad": grad_err((k, p.grad) for k, p in model.named_parameters())}
    ref = {"loss": abs(float(ytotal) - ototal.item()) / abs(ototal.item()),
           "grad": grad_err((k, v.grad) for k, v in sdy.items() if v.is_floating_point())}
    for k in ours:
        assert ours[k] <= max(3.0 * ref[k], tol), (k, ours, ref)
    # the local decoder feeds no loss term (train_mtmm_sd.py:240-252 uses g_depth_out only): no gradient reaches it
    assert all(p.grad is None or float(p.grad.abs().max()) == 0.0 for p in model.local_decoder.parameters())
