"""Generate golden fixtures by running the LIVE reference (/root/reference) on CPU.

    python tests/golden/make_golden.py          (build container only; /root/reference must exist)

Writes small .npz files next to this script.  They travel to the GPU box, where /root/reference does
not exist, and pin the oracle (tests/test_oracle_golden.py) and the CUDA path (tests/test_*_gpu.py).
Weights and inputs are produced by oracle.ref_oracle.build_tsn_state / synthetic_clip_batch (numpy
RandomState: machine independent), loaded into the reference with strict=True.
"""
from __future__ import annotations

import ast
import contextlib
import io
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REF))

from oracle import ref_oracle as O  # noqa: E402


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def ref_functions(path: Path, names):
    """Extract top-level function / class definitions from a reference script WITHOUT importing it (the train
    scripts parse argv at import time) and compile them in a scratch namespace."""
    from copy import deepcopy
    tree = ast.parse(path.read_text())
    keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    ns = {"torch": torch, "F": F, "deepcopy": deepcopy}
    exec(compile(ast.Module(body=keep, type_ignores=[]), str(path), "exec"), ns)
    return [ns[n] for n in names]


def grad_digest(named_grads):
    """Per-parameter (sum, sum of |g|, g.flat[::stride] samples) — small but discriminating."""
    out = {}
    for k, g in named_grads.items():
        g = g.detach().double().flatten()
        idx = torch.linspace(0, g.numel() - 1, steps=min(16, g.numel())).long()
        out[k] = np.concatenate([[g.sum().item(), g.abs().sum().item()], g[idx].numpy()])
    return out


def golden_shift():
    from models.temporal_shift import InplaceShift, TemporalShift
    rs = np.random.RandomState(7)
    cases = {}
    for name, (nt, c, h, w, T, div) in {
        "a": (8, 16, 5, 5, 4, 8), "b": (16, 24, 7, 7, 8, 8), "c": (6, 3, 4, 4, 3, 8), "d": (8, 20, 3, 3, 8, 3),
        "e": (16, 160, 7, 7, 8, 8), "f": (4, 8, 2, 2, 1, 4), "g": (8, 9, 3, 5, 2, 2),
    }.items():
        x = torch.from_numpy(rs.standard_normal((nt, c, h, w)).astype(np.float32)).requires_grad_(True)
        y = TemporalShift.shift(x, T, fold_div=div)
        g = torch.from_numpy(rs.standard_normal((nt, c, h, w)).astype(np.float32))
        y.backward(g)
        # the reference's in-place variant (unreachable via shift(); called directly) must agree
        xin = x.detach().clone().view(nt // T, T, c, h, w)
        yin = InplaceShift.apply(xin, c // div)
        assert torch.equal(yin.reshape(nt, c, h, w), y.detach())
        cases[name + "_meta"] = np.array([nt, c, h, w, T, div])
        cases[name + "_x"] = x.detach().numpy()
        cases[name + "_y"] = y.detach().numpy()
        cases[name + "_g"] = g.numpy()
        cases[name + "_gx"] = x.grad.numpy()
    np.savez_compressed(HERE / "shift.npz", **cases)


def golden_action():
    from models.action import Action
    out = {}
    for name, (c, h, T, n, train_bn) in {"c32": (32, 6, 4, 2, True), "c24": (24, 5, 8, 1, True),
                                        "c160": (160, 3, 8, 1, False)}.items():
        rs = np.random.RandomState(11 + c)
        sd = {}
        O.action_state(sd, "m", c, 8, rs)
        O._conv_entry(sd, "m.net.weight", (6 * c, c, 1, 1), rs)
        with _quiet():
            mod = Action(torch.nn.Conv2d(c, 6 * c, 1, bias=False), n_segment=T, shift_div=8)
        mod.load_state_dict({k[2:]: v for k, v in sd.items()}, strict=True)
        mod.train(train_bn)
        x = torch.from_numpy(rs.standard_normal((n * T, c, h, h)).astype(np.float32)).requires_grad_(True)
        y = mod(x)
        g = torch.from_numpy(rs.standard_normal(tuple(y.shape)).astype(np.float32))
        y.backward(g)
        out[name + "_meta"] = np.array([c, h, T, n, int(train_bn)])
        out[name + "_x"] = x.detach().numpy()
        out[name + "_g"] = g.numpy()
        out[name + "_y"] = y.detach().numpy()
        out[name + "_gx"] = x.grad.numpy()
        for k, p in mod.named_parameters():
            out[f"{name}_grad_{k}"] = p.grad.numpy()
        out[name + "_rm"] = mod.action_p3_bn1.running_mean.numpy()
        out[name + "_rv"] = mod.action_p3_bn1.running_var.numpy()
    np.savez_compressed(HERE / "action.npz", **out)


TSN_FIXTURE_CLIPS, TSN_FIXTURE_SIZE = 4, 96


def golden_tsn():
    """Whole TSN-MobileNetV2 fwd+bwd at 96x96, 4 clips (final map 3x3: the last BatchNorms see 288 samples per
    channel — the round-1 fixture, 2 clips at 64x64, gave them 64 and was so ill-conditioned that the reference's own
    fp32 run sat 5e-3 from its fp64 run), the three temporal variants, train-BN and frozen-BN (eval) modes."""
    from models.models import TSN
    from models.temporal_shift import TemporalShift
    from archs.mobilenet_v2 import InvertedResidual
    out = {}
    for temporal in ("none", "tsm", "action"):
        for bn_train in (True, False):
            for dt, dtag in ((torch.float32, ""), (torch.float64, "64")):
                with _quiet():
                    ref = TSN(83, 8, 'RGB', base_model='mobilenetv2', pretrain=None, dropout=0.5, partial_bn=False,
                              is_shift=(temporal == "action"), shift_div=8, consensus_type='avg', fc_lr5=True,
                              img_feature_dim=224)
                    if temporal == "tsm":
                        for m in ref.base_model.modules():
                            if isinstance(m, InvertedResidual) and len(m.conv) == 8 and m.use_res_connect:
                                m.conv[0] = TemporalShift(m.conv[0], n_segment=8, n_div=8)
                sd = O.build_tsn_state(83, temporal, 8, seed=5)
                ref.load_state_dict(sd, strict=True)       # checkpoint-contract check
                ref = ref.to(dt)
                ref.train(bn_train)
                for m in ref.modules():
                    if isinstance(m, torch.nn.Dropout):
                        m.eval()
                rgb, _, labels = O.synthetic_clip_batch(TSN_FIXTURE_CLIPS, 8, TSN_FIXTURE_SIZE, 83, seed=3)
                logits = ref(rgb.to(dt))
                loss = F.cross_entropy(logits, labels)
                loss.backward()
                tag = f"{temporal}_{'train' if bn_train else 'eval'}"
                out[tag + f"_logits{dtag}"] = logits.detach().numpy()
                out[tag + f"_loss{dtag}"] = np.array(loss.item())
                for k, v in grad_digest({k: p.grad for k, p in ref.named_parameters()}).items():
                    out[f"{tag}_g{dtag}_{k}"] = v
                # full gradient tensors of the first and the last layer (the digests above sample 16 entries each)
                named = dict(ref.named_parameters())
                out[f"{tag}_gfull{dtag}_base_model.features.0.0.weight"] = named["base_model.features.0.0.weight"].grad.numpy().copy()
                out[f"{tag}_gfull{dtag}_new_fc.bias"] = named["new_fc.bias"].grad.numpy().copy()
                out[f"{tag}_gfull{dtag}_new_fc.weight"] = named["new_fc.weight"].grad[:8].numpy().copy()
                rsd = ref.state_dict()
                for k in ("base_model.features.0.1.running_mean", "base_model.features.9.conv.4.running_var",
                          "base_model.features.18.1.running_var"):
                    out[f"{tag}_rs{dtag}_{k}"] = rsd[k].numpy()
    np.savez_compressed(HERE / "tsn_mbv2.npz", **out)


def golden_losses():
    kd_fn, feat_fn = ref_functions(REF / "train_sd.py", ["kd_loss_function", "feature_loss_function"])

    class A:  # the argparse namespace the reference functions read
        temperature, alpha, beta = 3, 0.1, 1e-6

    rs = np.random.RandomState(21)
    N, T, cls, fd = 4, 8, 83, 1280
    logits = [torch.from_numpy(rs.standard_normal((N, cls)).astype(np.float32) * 2).requires_grad_(True) for _ in range(4)]
    feats = [torch.from_numpy(rs.standard_normal((N * T, fd, 1, 1)).astype(np.float32)).requires_grad_(True) for _ in range(4)]
    labels = torch.from_numpy(rs.randint(0, cls, (N,)).astype(np.int64))
    crit = torch.nn.CrossEntropyLoss()
    # train_sd.py:227-265, statement for statement
    output, m1, m2, m3 = logits
    final_fea, f1, f2, f3 = feats
    loss = crit(output, labels)
    ml = [crit(m, labels) for m in (m1, m2, m3)]
    temp4 = torch.softmax(output / A.temperature, dim=1)
    kd = [kd_fn(m, temp4.detach(), A) * (A.temperature ** 2) for m in (m1, m2, m3)]
    fl = [feat_fn(f, final_fea.detach()) for f in (f1, f2, f3)]
    total = (1 - A.alpha) * (loss + ml[0] + ml[1] + ml[2]) + A.alpha * (kd[0] + kd[1] + kd[2]) + A.beta * (fl[0] + fl[1] + fl[2])
    total.backward()
    out = {"sd_labels": labels.numpy(), "sd_total": np.array(total.item()),
           "sd_terms": np.array([loss.item()] + [v.item() for v in ml] + [v.item() for v in kd] + [v.item() for v in fl])}
    for i in range(4):
        out[f"sd_logits{i}"] = logits[i].detach().numpy()
        out[f"sd_feat{i}"] = feats[i].detach().numpy()
        out[f"sd_glogits{i}"] = logits[i].grad.numpy()
        out[f"sd_gfeat{i}"] = feats[i].grad.numpy() if feats[i].grad is not None else np.zeros_like(feats[i].detach().numpy())

    # MTMM loss, train_mtmm.py:223-231
    lg = torch.from_numpy(rs.standard_normal((N, cls)).astype(np.float32)).requires_grad_(True)
    pred = torch.from_numpy(rs.uniform(0, 1, (N * 2, 1, 56, 56)).astype(np.float32)).requires_grad_(True)
    n_depth = torch.from_numpy(rs.uniform(0, 1, (N, 2, 1, 224, 224)).astype(np.float32))
    gt = F.interpolate(n_depth.view(-1, 1, n_depth.size(-2), n_depth.size(-1)), size=(56, 56), mode='bilinear')
    g_depth_loss = torch.nn.MSELoss()(pred, gt)
    mt = crit(lg, labels) + 0.01 * g_depth_loss
    mt.backward()
    out.update({"mt_logits": lg.detach().numpy(), "mt_pred": pred.detach().numpy(),
                "mt_depth": n_depth.numpy().astype(np.float16).astype(np.float32), "mt_loss": np.array(mt.item()),
                "mt_depth_loss": np.array(g_depth_loss.item()), "mt_glogits": lg.grad.numpy(),
                "mt_gpred": pred.grad.numpy()})
    # depth is stored through fp16 to keep the fixture small: recompute the reference on that input
    n_depth = torch.from_numpy(out["mt_depth"])
    pred2 = pred.detach().clone().requires_grad_(True)
    lg2 = lg.detach().clone().requires_grad_(True)
    gt = F.interpolate(n_depth.view(-1, 1, 224, 224), size=(56, 56), mode='bilinear')
    g_depth_loss = torch.nn.MSELoss()(pred2, gt)
    mt = crit(lg2, labels) + 0.01 * g_depth_loss
    mt.backward()
    out.update({"mt_loss": np.array(mt.item()), "mt_depth_loss": np.array(g_depth_loss.item()),
                "mt_glogits": lg2.grad.numpy(), "mt_gpred": pred2.grad.numpy(),
                "mt_depth": out["mt_depth"].astype(np.float16)})
    np.savez_compressed(HERE / "losses.npz", **out)


def _fwd_bwd(mod, x, rs):
    x = x.clone().requires_grad_(True)
    y = mod(x)
    g = torch.from_numpy(rs.standard_normal(tuple(y.shape)).astype(np.float32))
    y.backward(g)
    return x, y, g


def golden_heads():
    """Pieces either side of the backbone that round 1 left unpinned: SepConv (models/models_SD.py:81-101), the MTMM
    depth decoder (models/models_MTMM.py:129-155, taken from a live models_MTMM.TSN), the MTMM+SD ConvTranspose
    decoders (models/models_MTMM_SD.py:226-249) and the combined loss (train_mtmm_sd.py:240-293).  Every case also
    ASSERTS here that the oracle restatement agrees with the live reference."""
    from models.models_SD import SepConv
    out = {}

    def check(a, b, tol=2e-5, what=""):
        err = (a.detach().double() - b.detach().double()).abs().max().item() / max(b.detach().abs().max().item(), 1e-30)
        assert err < tol, (what, err)

    # ---- SepConv --------------------------------------------------------------------------
    for name, (ci, co, h, n, train_bn) in {"sep_a": (24, 32, 9, 4, True), "sep_b": (96, 160, 6, 2, True),
                                           "sep_c": (32, 96, 7, 2, False)}.items():
        rs = np.random.RandomState(501 + ci)
        sd = {}
        O.sepconv_state(sd, "m", ci, co, rs)
        mod = SepConv(ci, co)
        mod.load_state_dict({k[2:]: v for k, v in sd.items()}, strict=True)
        mod.train(train_bn)
        x0 = torch.from_numpy(rs.standard_normal((n, ci, h, h)).astype(np.float32))
        x, y, g = _fwd_bwd(mod, x0, rs)
        osd = O.clone_state(sd)
        xo = x0.clone().requires_grad_(True)
        yo = O.sepconv(xo, osd, "m", train_bn)
        yo.backward(g)
        check(yo, y, what=name)
        check(xo.grad, x.grad, what=name + " gx")
        out[name + "_meta"] = np.array([ci, co, h, n, int(train_bn)])
        out[name + "_x"], out[name + "_g"], out[name + "_y"], out[name + "_gx"] = (x0.numpy(), g.numpy(), y.detach().numpy(),
                                                                               x.grad.numpy())
        for k, p_ in mod.named_parameters():
            check(osd["m." + k].grad, p_.grad, what=name + k)
            out[f"{name}_grad_{k}"] = p_.grad.numpy()
        for k, b in mod.named_buffers():
            if "running" in k:
                out[f"{name}_buf_{k}"] = b.numpy().copy()

    # ---- MTMM global_decoder: the module of a live models_MTMM.TSN (ResNet-50: 2048 input channels) --------
    from models.models_MTMM import TSN as TSN_MTMM
    with _quiet():
        ref = TSN_MTMM(83, 8, 'RGB', base_model='resnet50', pretrain=None, dropout=0.5, partial_bn=False, is_shift=True,
                       shift_div=8, consensus_type='avg', fc_lr5=True, img_feature_dim=224, modal='rgb_depth')
    dec = ref.global_decoder
    for name, (h, n, train_bn) in {"dec_a": (3, 2, True), "dec_b": (2, 3, False)}.items():
        rs = np.random.RandomState(601 + h)
        sd = {}
        O.decoder_state(sd, rs, feat=2048, prefix="d")
        dec.load_state_dict({k[2:]: v for k, v in sd.items()}, strict=True)
        dec.train(train_bn)
        dec.zero_grad()
        x0 = torch.from_numpy(rs.standard_normal((n, 2048, h, h)).astype(np.float32))
        x, y, g = _fwd_bwd(dec, x0, rs)
        osd = O.clone_state(sd)
        xo = x0.clone().requires_grad_(True)
        yo = O.global_decoder(xo, osd, train_bn, prefix="d")
        yo.backward(g)
        check(yo, y, what=name)
        check(xo.grad, x.grad, what=name + " gx")
        out[name + "_meta"] = np.array([h, n, int(train_bn)])
        out[name + "_x"], out[name + "_g"], out[name + "_y"], out[name + "_gx"] = (x0.numpy().astype(np.float16), g.numpy(),
                                                                               None, None)
        # x is stored through fp16 (fixture size): recompute the reference on the rounded input
        x0 = torch.from_numpy(out[name + "_x"].astype(np.float32))
        dec.load_state_dict({k[2:]: v for k, v in sd.items()}, strict=True)
        dec.zero_grad()
        x = x0.clone().requires_grad_(True)
        y = dec(x)
        y.backward(g)
        out[name + "_y"], out[name + "_gx"] = y.detach().numpy(), x.grad.numpy()
        for k, p_ in dec.named_parameters():
            gk = p_.grad.detach().double().flatten()
            idx = torch.linspace(0, gk.numel() - 1, steps=min(64, gk.numel())).long()
            out[f"{name}_gdig_{k}"] = np.concatenate([[gk.sum().item(), gk.abs().sum().item()], gk[idx].numpy()])
        out[name + "_rv13"] = dec[13].running_var.numpy().copy()
    del ref

    # ---- MTMM+SD ConvTranspose decoders from a live models_MTMM_SD.TSN ------------------------------------
    from models.models_MTMM_SD import TSN as TSN_MS
    with _quiet():
        ref = TSN_MS(83, 8, 'RGB', base_model='resnet50', pretrain=None, dropout=0.5, partial_bn=False, is_shift=True,
                     shift_div=8, consensus_type='avg', fc_lr5=True, img_feature_dim=224, modal='rgb_depth')
    for name, (mod, chans, h, n, train_bn) in {"ctl": (ref.local_decoder, (64, 32, 1), 5, 2, True),
                                               "ctg": (ref.global_decoder, (2048, 256, 32, 1), 2, 2, True),
                                               "cte": (ref.local_decoder, (64, 32, 1), 4, 3, False)}.items():
        rs = np.random.RandomState(701 + h)
        sd = {}
        O.convt_decoder_state(sd, "d", chans, rs)
        mod.load_state_dict({k[2:]: v for k, v in sd.items()}, strict=True)
        mod.train(train_bn)
        mod.zero_grad()
        x0 = torch.from_numpy(rs.standard_normal((n, chans[0], h, h)).astype(np.float16).astype(np.float32))
        x, y, g = _fwd_bwd(mod, x0, rs)
        osd = O.clone_state(sd)
        xo = x0.clone().requires_grad_(True)
        yo = O.convt_decoder(xo, osd, "d", len(chans) - 1, train_bn)
        yo.backward(g)
        check(yo, y, what=name)
        check(xo.grad, x.grad, what=name + " gx")
        out[name + "_meta"] = np.array(list(chans) + [h, n, int(train_bn)])
        out[name + "_x"], out[name + "_g"], out[name + "_y"], out[name + "_gx"] = (x0.numpy().astype(np.float16), g.numpy(),
                                                                               y.detach().numpy(), x.grad.numpy())
        for k, p_ in mod.named_parameters():
            check(osd["d." + k].grad, p_.grad, what=name + k)
            gk = p_.grad.detach().double().flatten()
            idx = torch.linspace(0, gk.numel() - 1, steps=min(64, gk.numel())).long()
            out[f"{name}_gdig_{k}"] = np.concatenate([[gk.sum().item(), gk.abs().sum().item()], gk[idx].numpy()])
    del ref

    # ---- combined MTMM+SD loss, train_mtmm_sd.py:240-293 statement for statement ---------------------------
    kd_fn, feat_fn = ref_functions(REF / "train_mtmm_sd.py", ["kd_loss_function", "feature_loss_function"])

    class A:
        temperature, alpha, beta = 3, 0.1, 1e-6

    rs = np.random.RandomState(33)
    N, T, cls, fd = 3, 2, 83, 1280
    lg = [torch.from_numpy(rs.standard_normal((N, cls)).astype(np.float32) * 2).requires_grad_(True) for _ in range(4)]
    ft = [torch.from_numpy(rs.standard_normal((N * T, fd, 1, 1)).astype(np.float32)).requires_grad_(True) for _ in range(4)]
    labels = torch.from_numpy(rs.randint(0, cls, (N,)).astype(np.int64))
    g_depth_out = torch.from_numpy(rs.uniform(0, 1, (N * T, 1, 56, 56)).astype(np.float32)).requires_grad_(True)
    depth = torch.from_numpy(rs.uniform(0, 1, (N, T, 1, 224, 224)).astype(np.float16).astype(np.float32))
    criterion, mse_loss = torch.nn.CrossEntropyLoss(), torch.nn.MSELoss()
    output, middle_output1, middle_output2, middle_output3 = lg
    final_fea, middle1_fea, middle2_fea, middle3_fea = ft
    l_depth_gt = depth.view(-1, 1, depth.size(-2), depth.size(-1))
    g_depth_gt = F.interpolate(l_depth_gt, size=(56, 56), mode='bilinear')
    g_depth_loss = mse_loss(g_depth_out, g_depth_gt)
    loss = criterion(output, labels) + 0.01 * g_depth_loss
    middle1_loss = criterion(middle_output1, labels)
    middle2_loss = criterion(middle_output2, labels)
    middle3_loss = criterion(middle_output3, labels)
    temp4 = torch.softmax(output / A.temperature, dim=1)
    loss1by4 = kd_fn(middle_output1, temp4.detach(), A) * (A.temperature ** 2)
    loss2by4 = kd_fn(middle_output2, temp4.detach(), A) * (A.temperature ** 2)
    loss3by4 = kd_fn(middle_output3, temp4.detach(), A) * (A.temperature ** 2)
    feature_loss_1 = feat_fn(middle1_fea, final_fea.detach())
    feature_loss_2 = feat_fn(middle2_fea, final_fea.detach())
    feature_loss_3 = feat_fn(middle3_fea, final_fea.detach())
    total_loss = (1 - A.alpha) * (loss + middle1_loss + middle2_loss + middle3_loss) + \
        A.alpha * (loss1by4 + loss2by4 + loss3by4) + A.beta * (feature_loss_1 + feature_loss_2 + feature_loss_3)
    total_loss.backward()
    ot, ol = O.mtmm_sd_loss([t.detach() for t in lg], [t.detach() for t in ft], g_depth_out.detach(), depth, labels)
    assert abs(ot.item() - total_loss.item()) < 1e-5 and abs(ol.item() - loss.item()) < 1e-6
    out.update({"ms_labels": labels.numpy(), "ms_total": np.array(total_loss.item()), "ms_loss": np.array(loss.item()),
                "ms_depth": depth.numpy().astype(np.float16), "ms_gpred_in": g_depth_out.detach().numpy(),
                "ms_gpred": g_depth_out.grad.numpy()})
    for i in range(4):
        out[f"ms_logits{i}"], out[f"ms_feat{i}"] = lg[i].detach().numpy(), ft[i].detach().numpy()
        out[f"ms_glogits{i}"] = lg[i].grad.numpy()
        out[f"ms_gfeat{i}"] = ft[i].grad.numpy() if ft[i].grad is not None else np.zeros_like(ft[i].detach().numpy())

    np.savez_compressed(HERE / "heads.npz", **out)


RESNET_FIXTURE = dict(clips=2, T=4, size=64, num_class=10, seed=11, in_seed=12)


def golden_resnet():
    """N3: ResNet-50.  The reference has no ResNet source: models/models.py:108-117 instantiates torchvision.models.resnet50 and
    models/temporal_shift.py:101-146 wraps conv1 of every bottleneck.  Pinned here: (a) live torchvision resnet50 + the
    REFERENCE's make_temporal_shift (the TSM variant of config #5) — the four stage outputs and every gradient, fp64;
    (b) the live reference TSN(base_model='resnet50', is_shift=False) — logits and gradients, fp64."""
    import torchvision
    from models.models import TSN
    from models.temporal_shift import make_temporal_shift
    cfg = RESNET_FIXTURE
    rgb, _, labels = O.synthetic_clip_batch(cfg["clips"], cfg["T"], cfg["size"], cfg["num_class"], seed=cfg["in_seed"])
    x = rgb.view((-1, 3) + tuple(rgb.shape[-2:])).double()
    out = {}
    rs = np.random.RandomState(5)
    # (a) TSM backbone
    for bn_train in (True, False):
        tag = "tsm_" + ("train" if bn_train else "eval")
        sd = O.build_resnet_state(O.RESNET50_LAYERS, cfg["num_class"], "tsm", seed=cfg["seed"])
        with _quiet():
            net = torchvision.models.resnet50()
            make_temporal_shift(net, cfg["T"], n_div=8, place='blockres')
        bsd = {k[len("base_model."):]: v for k, v in sd.items() if k.startswith("base_model.")}
        missing = net.load_state_dict(bsd, strict=False)
        assert not missing.unexpected_keys and set(missing.missing_keys) == {"fc.weight", "fc.bias"}, missing
        net = net.double()
        net.train(bn_train)
        y = net.maxpool(net.relu(net.bn1(net.conv1(x))))
        t1 = net.layer1(y); t2 = net.layer2(t1); t3 = net.layer3(t2); t4 = net.layer4(t3)
        live = (t1, t2, t3, t4)
        if bn_train:
            out["gout_seed"] = np.array(5)
        gouts = [torch.from_numpy(np.random.RandomState(100 + i).standard_normal(tuple(t.shape))) / t.numel() ** 0.5
                 for i, t in enumerate(live)]
        torch.autograd.backward(live, gouts)
        osd = O.clone_state(sd, dtype=torch.float64)
        mine = O.resnet_features(x, osd, O.RESNET50_LAYERS, "tsm", cfg["T"], 8, bn_train)
        torch.autograd.backward(mine, gouts)
        for i, (a, b) in enumerate(zip(mine, live)):
            assert (a - b).abs().max().item() <= 1e-9 * b.abs().max().item(), (tag, i)
            out[f"{tag}_tap{i + 1}_chansum"] = b.detach().sum((0, 2, 3)).numpy()
            out[f"{tag}_tap{i + 1}_abs"] = np.array(b.detach().abs().sum().item())
        out[tag + "_layer4"] = live[3].detach().numpy()
        named = {"base_model." + k: p for k, p in net.named_parameters() if not k.startswith("fc.")}
        gmax = max(p.grad.abs().max().item() for p in named.values())
        for k, p in named.items():
            assert (osd[k].grad - p.grad).abs().max().item() <= 1e-9 * gmax, (tag, k)
        for k, v in grad_digest({k: p.grad for k, p in named.items()}).items():
            out[f"{tag}_g_{k}"] = v
        out[f"{tag}_gfull_base_model.conv1.weight"] = named["base_model.conv1.weight"].grad.numpy().copy()
        rsd = net.state_dict()
        for k in ("bn1.running_mean", "layer2.0.downsample.1.running_var", "layer4.2.bn3.running_var"):
            assert (osd["base_model." + k] - rsd[k]).abs().max().item() <= 1e-12 * rsd[k].abs().max().item()
            out[f"{tag}_rs_base_model.{k}"] = rsd[k].numpy()
    # (b) the reference's own TSN wrapper on resnet50, no temporal module
    sd = O.build_resnet_state(O.RESNET50_LAYERS, cfg["num_class"], "none", seed=cfg["seed"])
    with _quiet():
        ref = TSN(cfg["num_class"], cfg["T"], 'RGB', base_model='resnet50', pretrain=None, dropout=0.5, partial_bn=False,
                  is_shift=False, consensus_type='avg', fc_lr5=True, img_feature_dim=224)
    ref.load_state_dict(sd, strict=True)
    ref = ref.double()
    ref.train()                                   # the reference's TSN.train() returns None
    for m in ref.modules():
        if isinstance(m, torch.nn.Dropout):
            m.eval()
    logits = ref(rgb.double())
    loss = F.cross_entropy(logits, labels)
    loss.backward()
    osd = O.clone_state(sd, dtype=torch.float64)
    ologits = O.resnet_tsn_forward(rgb.double(), osd, cfg["T"], O.RESNET50_LAYERS, "none", 8, True)
    assert (ologits - logits).abs().max().item() <= 1e-10 * logits.abs().max().item()
    F.cross_entropy(ologits, labels).backward()
    gmax = max(p.grad.abs().max().item() for p in ref.parameters())
    for k, p in ref.named_parameters():
        assert (osd[k].grad - p.grad).abs().max().item() <= 1e-9 * gmax, k
    out["tsn_none_logits"] = logits.detach().numpy()
    out["tsn_none_loss"] = np.array(loss.item())
    for k, v in grad_digest({k: p.grad for k, p in ref.named_parameters()}).items():
        out[f"tsn_none_g_{k}"] = v
    np.savez_compressed(HERE / "resnet.npz", **out)


def golden_resnet_wrappers():
    """N3: the UNMODIFIED reference wrappers on their own backbone — models_MTMM.TSN and models_SD.TSN with
    base_model='resnet50' (is_shift=False) — against the oracle's resnet_mtmm_forward / resnet_sd_forward in fp64: outputs and
    every gradient.  Written into resnet_wrappers.npz (outputs + gradient digests)."""
    import models.models_MTMM as MT
    import models.models_SD as SDM
    cfg = RESNET_FIXTURE
    rgb, depth, labels = O.synthetic_clip_batch(cfg["clips"], cfg["T"], cfg["size"], cfg["num_class"], seed=cfg["in_seed"])
    out = {}
    common = dict(base_model='resnet50', pretrain=None, dropout=0.5, partial_bn=False, is_shift=False, consensus_type='avg',
                  fc_lr5=True, img_feature_dim=224)

    def load(ref, sd):
        res = ref.load_state_dict(sd, strict=False)
        assert not res.unexpected_keys, res.unexpected_keys
        # models_MTMM keeps a torch.fx feature extractor over the SAME parameters (models/models_MTMM.py:70-77)
        assert all(k.startswith("feature_extractor.") for k in res.missing_keys), res.missing_keys
        ref.double()
        ref.train()
        for m in ref.modules():
            if isinstance(m, torch.nn.Dropout):
                m.eval()

    # ---- MTMM
    sd = O.build_resnet_mtmm_state(cfg["num_class"], "none", seed=cfg["seed"])
    with _quiet():
        ref = MT.TSN(cfg["num_class"], cfg["T"], 'RGB', modal='rgb_depth', **common)
    load(ref, sd)
    logits, dpred = ref(rgb.double())
    gt = F.interpolate(depth.double().view(-1, 1, cfg["size"], cfg["size"]), tuple(dpred.shape[-2:]), mode='bilinear')
    (F.cross_entropy(logits, labels) + 0.01 * F.mse_loss(dpred, gt)).backward()
    osd = O.clone_state(sd, dtype=torch.float64)
    ol, od = O.resnet_mtmm_forward(rgb.double(), osd, cfg["T"], "none", 8, True)
    (F.cross_entropy(ol, labels) + 0.01 * F.mse_loss(od, gt)).backward()
    assert (ol - logits).abs().max().item() <= 1e-10 * logits.abs().max().item()
    assert (od - dpred).abs().max().item() <= 1e-10
    named = {k: p for k, p in ref.named_parameters() if not k.startswith("feature_extractor.")}
    gmax = max(p.grad.abs().max().item() for p in named.values())
    for k, p in named.items():
        assert (osd[k].grad - p.grad).abs().max().item() <= 1e-9 * gmax, k
    out["mtmm_logits"], out["mtmm_depth"] = logits.detach().numpy(), dpred.detach().numpy()
    for k, v in grad_digest({k: p.grad for k, p in named.items()}).items():
        out[f"mtmm_g_{k}"] = v
    # ---- SD
    sd = O.build_resnet_sd_state(cfg["num_class"], "none", seed=cfg["seed"])
    with _quiet():
        ref = SDM.TSN(cfg["num_class"], cfg["T"], 'RGB', **common)
    res = ref.load_state_dict(sd, strict=True)
    ref.double()
    ref.train()
    for m in ref.modules():
        if isinstance(m, torch.nn.Dropout):
            m.eval()
    outs = ref(rgb.double())
    (kd_loss_function, feature_loss_function) = ref_functions(REF / "train_sd.py", ["kd_loss_function", "feature_loss_function"])
    class _A: temperature = 3.0
    def total_of(o):
        ce = sum(F.cross_entropy(z, labels) for z in o[:4])
        temp4 = torch.softmax(o[0] / _A.temperature, dim=1)                # train_sd.py:236-237
        kd = sum(kd_loss_function(z, temp4.detach(), _A) * 9.0 for z in o[1:4])
        fe = sum(feature_loss_function(f, o[4].detach()) for f in o[5:8])
        return 0.9 * ce + 0.1 * kd + 1e-6 * fe
    total_of(outs).backward()
    osd = O.clone_state(sd, dtype=torch.float64)
    oouts = O.resnet_sd_forward(rgb.double(), osd, cfg["T"], "none", 8, True)
    ototal, _ = O.sd_loss(oouts[:4], oouts[4:], labels, 0.1, 1e-6, 3.0)
    ototal.backward()
    for i, (a, b) in enumerate(zip(oouts, outs)):
        assert a.shape == b.shape and (a - b).abs().max().item() <= 1e-10 * max(b.abs().max().item(), 1e-30), i
        out[f"sd_out{i}"] = b.detach().numpy()
    gmax = max(p.grad.abs().max().item() for p in ref.parameters())
    for k, p in ref.named_parameters():
        assert (osd[k].grad - p.grad).abs().max().item() <= 1e-9 * gmax, k
    for k, v in grad_digest({k: p.grad for k, p in ref.named_parameters()}).items():
        out[f"sd_g_{k}"] = v
    # ---- MTMM+SD (models/models_MTMM_SD.py: ResNet only; two backbone passes there, one in the oracle)
    import models.models_MTMM_SD as MS
    sd = O.build_resnet_mtmm_sd_state(cfg["num_class"], "none", seed=cfg["seed"])
    with _quiet():
        ref = MS.TSN(cfg["num_class"], cfg["T"], 'RGB', modal='rgb_depth', **common)
    load(ref, sd)
    outs = ref(rgb.double())
    gt = F.interpolate(depth.double().view(-1, 1, cfg["size"], cfg["size"]), tuple(outs[9].shape[-2:]), mode='bilinear')

    def combined(o):            # train_mtmm_sd.py:240-293 (+ a term on local_depth_out so that local_decoder gets a gradient)
        loss = F.cross_entropy(o[0], labels) + 0.01 * F.mse_loss(o[9], gt)
        ce = sum(F.cross_entropy(z, labels) for z in o[1:4])
        temp4 = torch.softmax(o[0] / _A.temperature, dim=1)
        kd = sum(kd_loss_function(z, temp4.detach(), _A) * 9.0 for z in o[1:4])
        fe = sum(feature_loss_function(f, o[4].detach()) for f in o[5:8])
        return 0.9 * (loss + ce) + 0.1 * kd + 1e-6 * fe + 0.01 * (o[8] ** 2).mean()
    combined(outs).backward()
    osd = O.clone_state(sd, dtype=torch.float64)
    oouts = O.resnet_mtmm_sd_forward(rgb.double(), osd, cfg["T"], "none", 8, True)
    combined(oouts).backward()
    assert len(outs) == len(oouts) == 10
    for i, (a, b) in enumerate(zip(oouts, outs)):
        assert a.shape == b.shape and (a - b).abs().max().item() <= 1e-10 * max(b.abs().max().item(), 1e-30), i
        if i not in (5, 6, 7):
            out[f"mtmmsd_out{i}"] = b.detach().numpy()
    named = {k: p for k, p in ref.named_parameters() if not k.startswith("feature_extractor.") and p.grad is not None}
    gmax = max(p.grad.abs().max().item() for p in named.values())
    for k, p in named.items():
        assert (osd[k].grad - p.grad).abs().max().item() <= 1e-9 * gmax, k
    for k, v in grad_digest({k: p.grad for k, p in named.items()}).items():
        out[f"mtmmsd_g_{k}"] = v
    np.savez_compressed(HERE / "resnet_wrappers.npz", **out)


def golden_ema_and_pool():
    """Separate small file: EMAWrapper replay (initial state stored explicitly) and TemporalPool."""
    from models.temporal_shift import TemporalPool
    (EMAWrapper,) = ref_functions(REF / "train_mtmm.py", ["EMAWrapper"])
    out = {}
    rs = np.random.RandomState(44)
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, bias=False), torch.nn.BatchNorm2d(4), torch.nn.Linear(4, 2))
    for k, v in net.state_dict().items():
        out["ema_init_" + k] = v.clone().numpy()
    ema = EMAWrapper(net, decay=0.9)
    oema = {k: v.clone() for k, v in net.state_dict().items()}
    for step in range(3):
        with torch.no_grad():
            for k, v in net.state_dict().items():
                if v.is_floating_point():
                    v.add_(torch.from_numpy(rs.standard_normal(tuple(v.shape)).astype(np.float32)))
                else:
                    v.add_(7 * (step + 1))
        for k, v in net.state_dict().items():
            out[f"ema_model{step}_{k}"] = v.clone().numpy()
        ema.update(net)
        O.ema_update(oema, net.state_dict(), 0.9)
    for k, v in ema.state_dict().items():
        assert torch.equal(v, oema[k]), k
        out["ema_final_" + k] = v.numpy()
    for name, (nt, c, h, T) in {"tp_a": (8, 5, 3, 8), "tp_b": (12, 4, 2, 4), "tp_c": (6, 3, 2, 6)}.items():
        x = torch.from_numpy(rs.standard_normal((nt, c, h, h)).astype(np.float32))
        y = TemporalPool.temporal_pool(x, T)
        assert np.array_equal(O.temporal_pool_np(x.numpy(), T), y.numpy())
        xg = x.clone().requires_grad_(True)
        TemporalPool.temporal_pool(xg, T).backward(torch.ones_like(y) * 0.5 + y.detach())
        out[name + "_meta"] = np.array([nt, c, h, T])
        out[name + "_x"], out[name + "_y"], out[name + "_gx"] = x.numpy(), y.numpy(), xg.grad.numpy()
    np.savez_compressed(HERE / "ema_pool.npz", **out)


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    only = set(sys.argv[1:])
    for name, fn in (("shift", golden_shift), ("action", golden_action), ("losses", golden_losses), ("tsn", golden_tsn),
                     ("heads", golden_heads), ("ema_pool", golden_ema_and_pool),
                     ("resnet", golden_resnet), ("resnet_wrappers", golden_resnet_wrappers)):
        if not only or name in only:
            fn()
    for p in sorted(HERE.glob("*.npz")):
        print(p.name, p.stat().st_size)
