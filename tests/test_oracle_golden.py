"""Pin the oracle against fixtures produced by the live reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_oracle as O
from conftest import GOLDEN, rel_err


def test_shift_oracle_matches_reference_bit_exact():
    z = np.load(GOLDEN / "shift.npz")
    names = sorted({k.split("_")[0] for k in z.files})
    assert len(names) >= 7
    for n in names:
        nt, c, h, w, T, div = z[n + "_meta"]
        assert np.array_equal(O.temporal_shift_np(z[n + "_x"], T, div), z[n + "_y"])
        assert np.array_equal(O.temporal_shift_bwd_np(z[n + "_g"], T, div), z[n + "_gx"])
        xt = torch.from_numpy(z[n + "_x"]).requires_grad_(True)
        yt = O.temporal_shift(xt, int(T), int(div))
        assert torch.equal(yt, torch.from_numpy(z[n + "_y"]))
        yt.backward(torch.from_numpy(z[n + "_g"]))
        assert torch.equal(xt.grad, torch.from_numpy(z[n + "_gx"]))


def test_shift_bad_segment_count_raises_like_reference():
    with pytest.raises(Exception):
        O.temporal_shift_np(np.zeros((7, 8, 2, 2), np.float32), 4, 8)


@pytest.mark.parametrize("name", ["c32", "c24", "c160"])
def test_action_oracle_matches_reference(name):
    z = np.load(GOLDEN / "action.npz")
    c, h, T, n, train_bn = z[name + "_meta"]
    rs = np.random.RandomState(11 + int(c))
    sd = {}
    O.action_state(sd, "m", int(c), 8, rs)
    O._conv_entry(sd, "m.net.weight", (6 * int(c), int(c), 1, 1), rs)
    sd = O.clone_state(sd)
    x = torch.from_numpy(z[name + "_x"]).requires_grad_(True)
    y = O.action_forward(x, sd, "m", int(T), bool(train_bn))
    assert rel_err(y, torch.from_numpy(z[name + "_y"])) < 2e-6
    y.backward(torch.from_numpy(z[name + "_g"]))
    assert rel_err(x.grad, torch.from_numpy(z[name + "_gx"])) < 1e-5
    for k in z.files:
        if k.startswith(name + "_grad_"):
            pk = "m." + k[len(name + "_grad_"):]
            ref = torch.from_numpy(z[k])
            if ref.abs().max() < 1e-7:
                assert sd[pk].grad.abs().max() < 1e-6
            else:
                assert rel_err(sd[pk].grad, ref) < 2e-5, pk
    assert rel_err(sd["m.action_p3_bn1.running_mean"], torch.from_numpy(z[name + "_rm"])) < 1e-5


def _digest(g):
    g = g.detach().double().flatten()
    idx = torch.linspace(0, g.numel() - 1, steps=min(16, g.numel())).long()
    return np.concatenate([[g.sum().item(), g.abs().sum().item()], g[idx].numpy()])


def _digest_err(sd, z, prefix):
    """max over parameters of |sampled grad - golden| / (largest sampled |grad| in the model)."""
    scale = max(np.abs(z[k][2:]).max() for k in z.files if k.startswith(prefix))
    worst = 0.0
    for k in z.files:
        if k.startswith(prefix):
            got = _digest(sd[k[len(prefix):]].grad)
            worst = max(worst, np.abs(got[2:] - z[k][2:]).max() / scale)
    return worst


@pytest.mark.parametrize("temporal", ["none", "tsm", "action"])
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_tsn_oracle_matches_reference(temporal, mode):
    """The oracle is pinned in fp64 (where op order does not matter: <=1e-9) and then checked in fp32
    against the reference's fp32 run with a bound that accounts for the conditioning of the problem
    (the reference's own fp32-vs-fp64 gradient error on this case is ~5e-3)."""
    z = np.load(GOLDEN / "tsn_mbv2.npz")
    tag = f"{temporal}_{mode}"
    rgb, _, labels = O.synthetic_clip_batch(4, 8, 96, 83, seed=3)
    # --- fp64 pin ---
    sd = O.clone_state(O.build_tsn_state(83, temporal, 8, seed=5), dtype=torch.float64)
    logits = O.tsn_forward(rgb.double(), sd, 8, temporal, 8, bn_training=(mode == "train"))
    assert rel_err(logits, torch.from_numpy(z[tag + "_logits64"])) < 1e-10
    F.cross_entropy(logits, labels).backward()
    assert _digest_err(sd, z, tag + "_g64_") < 1e-9
    for k in z.files:
        if k.startswith(tag + "_rs64_"):
            assert rel_err(sd[k[len(tag + "_rs64_"):]], torch.from_numpy(z[k])) < 1e-12
    # --- fp32 ---
    sd = O.clone_state(O.build_tsn_state(83, temporal, 8, seed=5))
    logits = O.tsn_forward(rgb, sd, 8, temporal, 8, bn_training=(mode == "train"))
    assert rel_err(logits, torch.from_numpy(z[tag + "_logits"])) < 1e-5
    loss = F.cross_entropy(logits, labels)
    assert abs(loss.item() - float(z[tag + "_loss"])) < 1e-5
    loss.backward()
    ref32_vs_64 = max(np.abs(z[k][2:] - z[k.replace("_g_", "_g64_")][2:]).max() for k in z.files if k.startswith(tag + "_g_")) \
        / max(np.abs(z[k][2:]).max() for k in z.files if k.startswith(tag + "_g64_"))
    assert _digest_err(sd, z, tag + "_g64_") < max(3 * ref32_vs_64, 1e-5)
    for k in z.files:
        if k.startswith(tag + "_rs_"):
            assert rel_err(sd[k[len(tag + "_rs_"):]], torch.from_numpy(z[k])) < 1e-5


def test_loss_oracles_match_reference():
    z = np.load(GOLDEN / "losses.npz")
    labels = torch.from_numpy(z["sd_labels"])
    logits = [torch.from_numpy(z[f"sd_logits{i}"]).requires_grad_(True) for i in range(4)]
    feats = [torch.from_numpy(z[f"sd_feat{i}"]).requires_grad_(True) for i in range(4)]
    total, terms = O.sd_loss(logits, feats, labels, 0.1, 1e-6, 3.0)
    assert abs(total.item() - float(z["sd_total"])) < 1e-5 * abs(float(z["sd_total"]))
    flat = [t.item() for t in terms["ce"]] + [t.item() for t in terms["kd"]] + [t.item() for t in terms["feat"]]
    assert np.allclose(flat, z["sd_terms"], rtol=1e-5)
    total.backward()
    for i in range(4):
        assert rel_err(logits[i].grad, torch.from_numpy(z[f"sd_glogits{i}"])) < 1e-5
        if i > 0:
            assert rel_err(feats[i].grad, torch.from_numpy(z[f"sd_gfeat{i}"])) < 1e-5
    assert feats[0].grad is None  # final features are detached (train_sd.py:252-259)

    lg = torch.from_numpy(z["mt_logits"]).requires_grad_(True)
    pred = torch.from_numpy(z["mt_pred"]).requires_grad_(True)
    depth = torch.from_numpy(z["mt_depth"].astype(np.float32))
    loss, dl = O.mtmm_loss(lg, labels, pred, depth)
    assert abs(loss.item() - float(z["mt_loss"])) < 1e-6 and abs(dl.item() - float(z["mt_depth_loss"])) < 1e-6
    loss.backward()
    assert rel_err(lg.grad, torch.from_numpy(z["mt_glogits"])) < 1e-6
    assert rel_err(pred.grad, torch.from_numpy(z["mt_gpred"])) < 1e-6


# ------------------------------------------------------------------------------------------------
# round 2: the pieces either side of the backbone (tests/golden/heads.npz, ema_pool.npz)
# ------------------------------------------------------------------------------------------------
def _digest64(g):
    g = g.detach().double().flatten()
    idx = torch.linspace(0, g.numel() - 1, steps=min(64, g.numel())).long()
    return np.concatenate([[g.sum().item(), g.abs().sum().item()], g[idx].numpy()])


@pytest.mark.parametrize("name", ["sep_a", "sep_b", "sep_c"])
def test_sepconv_oracle_matches_reference(name):
    """models/models_SD.py:81-101, fixture from the live reference SepConv."""
    z = np.load(GOLDEN / "heads.npz")
    ci, co, h, n, train_bn = (int(v) for v in z[name + "_meta"])
    sd = {}
    O.sepconv_state(sd, "m", ci, co, np.random.RandomState(501 + ci))
    sd = O.clone_state(sd)
    x = torch.from_numpy(z[name + "_x"]).requires_grad_(True)
    y = O.sepconv(x, sd, "m", bool(train_bn))
    assert rel_err(y, torch.from_numpy(z[name + "_y"])) < 2e-6
    y.backward(torch.from_numpy(z[name + "_g"]))
    assert rel_err(x.grad, torch.from_numpy(z[name + "_gx"])) < 1e-5
    for k in z.files:
        if k.startswith(name + "_grad_"):
            assert rel_err(sd["m." + k[len(name + "_grad_"):]].grad, torch.from_numpy(z[k])) < 2e-5, k
        if k.startswith(name + "_buf_") and train_bn:
            assert rel_err(sd["m." + k[len(name + "_buf_"):]], torch.from_numpy(z[k])) < 1e-5, k


@pytest.mark.parametrize("name", ["dec_a", "dec_b"])
def test_global_decoder_oracle_matches_reference(name):
    """models/models_MTMM.py:129-155, fixture from the global_decoder of a live models_MTMM.TSN (ResNet-50)."""
    z = np.load(GOLDEN / "heads.npz")
    h, n, train_bn = (int(v) for v in z[name + "_meta"])
    sd = {}
    O.decoder_state(sd, np.random.RandomState(601 + h), feat=2048, prefix="d")
    sd = O.clone_state(sd)
    x = torch.from_numpy(z[name + "_x"].astype(np.float32)).requires_grad_(True)
    y = O.global_decoder(x, sd, bool(train_bn), prefix="d")
    assert tuple(y.shape) == (n, 1, 8 * h, 8 * h)
    assert rel_err(y, torch.from_numpy(z[name + "_y"])) < 2e-6
    y.backward(torch.from_numpy(z[name + "_g"]))
    assert rel_err(x.grad, torch.from_numpy(z[name + "_gx"])) < 2e-5
    for k in z.files:
        if k.startswith(name + "_gdig_"):
            got, ref = _digest64(sd["d." + k[len(name + "_gdig_"):]].grad), z[k]
            assert np.abs(got - ref).max() <= 2e-5 * max(np.abs(ref[2:]).max(), 1e-6) + 1e-4 * abs(ref[1]) * 1e-3, k
    if train_bn:
        assert rel_err(sd["d.13.running_var"], torch.from_numpy(z[name + "_rv13"])) < 1e-5


@pytest.mark.parametrize("name", ["ctl", "ctg", "cte"])
def test_convt_decoder_oracle_matches_reference(name):
    """models/models_MTMM_SD.py:226-249: local / global ConvTranspose decoders of a live models_MTMM_SD.TSN."""
    z = np.load(GOLDEN / "heads.npz")
    meta = [int(v) for v in z[name + "_meta"]]
    chans, (h, n, train_bn) = meta[:-3], meta[-3:]
    sd = {}
    O.convt_decoder_state(sd, "d", chans, np.random.RandomState(701 + h))
    sd = O.clone_state(sd)
    x = torch.from_numpy(z[name + "_x"].astype(np.float32)).requires_grad_(True)
    y = O.convt_decoder(x, sd, "d", len(chans) - 1, bool(train_bn))
    assert rel_err(y, torch.from_numpy(z[name + "_y"])) < 2e-6
    y.backward(torch.from_numpy(z[name + "_g"]))
    assert rel_err(x.grad, torch.from_numpy(z[name + "_gx"])) < 2e-5
    for k in z.files:
        if k.startswith(name + "_gdig_"):
            got, ref = _digest64(sd["d." + k[len(name + "_gdig_"):]].grad), z[k]
            assert np.abs(got[2:] - ref[2:]).max() <= 3e-5 * max(np.abs(ref[2:]).max(), 1e-6), k


def test_mtmm_sd_loss_oracle_matches_reference():
    """train_mtmm_sd.py:240-293 (fixture = the reference statements with its own kd / feature functions)."""
    z = np.load(GOLDEN / "heads.npz")
    lg = [torch.from_numpy(z[f"ms_logits{i}"]).requires_grad_(True) for i in range(4)]
    ft = [torch.from_numpy(z[f"ms_feat{i}"]).requires_grad_(True) for i in range(4)]
    pred = torch.from_numpy(z["ms_gpred_in"]).requires_grad_(True)
    depth = torch.from_numpy(z["ms_depth"].astype(np.float32))
    total, loss = O.mtmm_sd_loss(lg, ft, pred, depth, torch.from_numpy(z["ms_labels"]))
    assert abs(total.item() - float(z["ms_total"])) < 1e-5 and abs(loss.item() - float(z["ms_loss"])) < 1e-6
    total.backward()
    for i in range(4):
        assert rel_err(lg[i].grad, torch.from_numpy(z[f"ms_glogits{i}"])) < 1e-5
        if i:
            assert rel_err(ft[i].grad, torch.from_numpy(z[f"ms_gfeat{i}"])) < 1e-5
    assert ft[0].grad is None or float(ft[0].grad.abs().max()) == 0.0
    assert rel_err(pred.grad, torch.from_numpy(z["ms_gpred"])) < 1e-5


def test_ema_oracle_matches_reference_bit_exact():
    """train_mtmm.py:110-128: three updates of the reference EMAWrapper (decay 0.9), replayed on the oracle."""
    z = np.load(GOLDEN / "ema_pool.npz")
    keys = [k[len("ema_init_"):] for k in z.files if k.startswith("ema_init_")]
    ema = {k: torch.from_numpy(z["ema_init_" + k].copy()) for k in keys}
    for step in range(3):
        O.ema_update(ema, {k: torch.from_numpy(z[f"ema_model{step}_{k}"]) for k in keys}, 0.9)
    for k in keys:
        assert torch.equal(ema[k], torch.from_numpy(z["ema_final_" + k])), k
    assert ema["1.num_batches_tracked"].dtype == torch.int64


@pytest.mark.parametrize("name", ["tp_a", "tp_b", "tp_c"])
def test_temporal_pool_oracle_matches_reference_bit_exact(name):
    z = np.load(GOLDEN / "ema_pool.npz")
    nt, c, h, T = (int(v) for v in z[name + "_meta"])
    assert np.array_equal(O.temporal_pool_np(z[name + "_x"], T), z[name + "_y"])


RESNET_FIXTURE = dict(clips=2, T=4, size=64, num_class=10, seed=11, in_seed=12)     # tests/golden/make_golden.py


def resnet_fixture_gouts(shapes):
    """The output gradients make_golden.golden_resnet used for the four stage outputs."""
    return [torch.from_numpy(np.random.RandomState(100 + i).standard_normal(tuple(s))) / float(np.prod(s)) ** 0.5
            for i, s in enumerate(shapes)]


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_resnet50_tsm_oracle_matches_live_torchvision_plus_reference_shift(mode):
    """N3: the reference instantiates torchvision.models.resnet50 (models/models.py:108-117) and wraps conv1 of every
    bottleneck (models/temporal_shift.py:101-146).  Fixture = that live pair in fp64; the oracle restatement must agree."""
    z = np.load(GOLDEN / "resnet.npz")
    cfg = RESNET_FIXTURE
    tag = "tsm_" + mode
    rgb, _, _ = O.synthetic_clip_batch(cfg["clips"], cfg["T"], cfg["size"], cfg["num_class"], seed=cfg["in_seed"])
    x = rgb.view((-1, 3) + tuple(rgb.shape[-2:])).double()
    sd = O.clone_state(O.build_resnet_state(O.RESNET50_LAYERS, cfg["num_class"], "tsm", seed=cfg["seed"]), dtype=torch.float64)
    taps = O.resnet_features(x, sd, O.RESNET50_LAYERS, "tsm", cfg["T"], 8, mode == "train")
    assert [tuple(t.shape[1:]) for t in taps] == [(256, 16, 16), (512, 8, 8), (1024, 4, 4), (2048, 2, 2)]
    for i, t in enumerate(taps):
        assert rel_err(t.sum((0, 2, 3)), torch.from_numpy(z[f"{tag}_tap{i + 1}_chansum"])) < 1e-9
        assert abs(t.abs().sum().item() - float(z[f"{tag}_tap{i + 1}_abs"])) < 1e-9 * float(z[f"{tag}_tap{i + 1}_abs"])
    assert rel_err(taps[3], torch.from_numpy(z[tag + "_layer4"])) < 1e-9
    torch.autograd.backward(taps, resnet_fixture_gouts([t.shape for t in taps]))
    assert _digest_err(sd, z, tag + "_g_") < 1e-9
    assert rel_err(sd["base_model.conv1.weight"].grad, torch.from_numpy(z[tag + "_gfull_base_model.conv1.weight"])) < 1e-9
    for k in z.files:
        if k.startswith(tag + "_rs_"):
            assert rel_err(sd[k[len(tag + "_rs_"):]], torch.from_numpy(z[k])) < 1e-12


def test_resnet50_tsn_oracle_matches_reference_wrapper():
    """The live reference TSN(base_model='resnet50', is_shift=False): logits, loss, gradient digests (fp64)."""
    z = np.load(GOLDEN / "resnet.npz")
    cfg = RESNET_FIXTURE
    rgb, _, labels = O.synthetic_clip_batch(cfg["clips"], cfg["T"], cfg["size"], cfg["num_class"], seed=cfg["in_seed"])
    sd = O.clone_state(O.build_resnet_state(O.RESNET50_LAYERS, cfg["num_class"], "none", seed=cfg["seed"]), dtype=torch.float64)
    logits = O.resnet_tsn_forward(rgb.double(), sd, cfg["T"], O.RESNET50_LAYERS, "none", 8, True)
    assert rel_err(logits, torch.from_numpy(z["tsn_none_logits"])) < 1e-10
    loss = F.cross_entropy(logits, labels)
    assert abs(loss.item() - float(z["tsn_none_loss"])) < 1e-10
    loss.backward()
    assert _digest_err(sd, z, "tsn_none_g_") < 1e-9


def test_resnet50_mtmm_and_sd_oracles_match_the_unmodified_reference_wrappers():
    """N3's purpose: the reference's OWN models_MTMM.TSN / models_SD.TSN run on ResNet bases.  Fixture = those
    live wrappers (base_model='resnet50', is_shift=False) in fp64: outputs and gradient digests of every parameter."""
    z = np.load(GOLDEN / "resnet_wrappers.npz")
    cfg = RESNET_FIXTURE
    rgb, depth, labels = O.synthetic_clip_batch(cfg["clips"], cfg["T"], cfg["size"], cfg["num_class"], seed=cfg["in_seed"])
    # MTMM: logits + depth map, loss of train_mtmm.py:223-231 at this resolution
    sd = O.clone_state(O.build_resnet_mtmm_state(cfg["num_class"], "none", seed=cfg["seed"]), dtype=torch.float64)
    ol, od = O.resnet_mtmm_forward(rgb.double(), sd, cfg["T"], "none", 8, True)
    assert rel_err(ol, torch.from_numpy(z["mtmm_logits"])) < 1e-10 and rel_err(od, torch.from_numpy(z["mtmm_depth"])) < 1e-10
    gt = F.interpolate(depth.double().view(-1, 1, cfg["size"], cfg["size"]), tuple(od.shape[-2:]), mode='bilinear')
    (F.cross_entropy(ol, labels) + 0.01 * F.mse_loss(od, gt)).backward()
    assert _digest_err(sd, z, "mtmm_g_") < 1e-9
    # SD: eight outputs, loss of train_sd.py:227-265
    sd = O.clone_state(O.build_resnet_sd_state(cfg["num_class"], "none", seed=cfg["seed"]), dtype=torch.float64)
    outs = O.resnet_sd_forward(rgb.double(), sd, cfg["T"], "none", 8, True)
    assert [tuple(o.shape) for o in outs] == [(2, 10)] * 4 + [(8, 2048, 1, 1)] * 4
    for i, o in enumerate(outs):
        assert rel_err(o, torch.from_numpy(z[f"sd_out{i}"])) < 1e-10, i
    total, _ = O.sd_loss(outs[:4], outs[4:], labels, 0.1, 1e-6, 3.0)
    total.backward()
    assert _digest_err(sd, z, "sd_g_") < 1e-9


def test_resnet50_mtmm_sd_oracle_matches_the_unmodified_reference_wrapper():
    """models_MTMM_SD.TSN (ResNet only in the reference, two backbone passes there): the oracle's single-pass restatement
    gives the same ten tensors and the same gradients of the combined loss (train_mtmm_sd.py:240-293), fp64."""
    z = np.load(GOLDEN / "resnet_wrappers.npz")
    cfg = RESNET_FIXTURE
    rgb, depth, labels = O.synthetic_clip_batch(cfg["clips"], cfg["T"], cfg["size"], cfg["num_class"], seed=cfg["in_seed"])
    sd = O.clone_state(O.build_resnet_mtmm_sd_state(cfg["num_class"], "none", seed=cfg["seed"]), dtype=torch.float64)
    o = O.resnet_mtmm_sd_forward(rgb.double(), sd, cfg["T"], "none", 8, True)
    assert len(o) == 10
    for i in (0, 1, 2, 3, 4, 8, 9):
        assert rel_err(o[i], torch.from_numpy(z[f"mtmmsd_out{i}"])) < 1e-10, i
    gt = F.interpolate(depth.double().view(-1, 1, cfg["size"], cfg["size"]), tuple(o[9].shape[-2:]), mode='bilinear')
    sd_total, _ = O.sd_loss(o[:4], o[4:8], labels, 0.1, 1e-6, 3.0)
    (sd_total + 0.9 * 0.01 * F.mse_loss(o[9], gt) + 0.01 * (o[8] ** 2).mean()).backward()
    assert _digest_err(sd, z, "mtmmsd_g_") < 1e-9
