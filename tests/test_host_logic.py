"""Host-side mirror of the reference interface: module tree, state_dict keys, optimiser groups,
error behaviour.  No GPU needed (nothing is executed on a device)."""
import contextlib
import io

import pytest
import torch

from oracle import ref_oracle as O


def _tsn(temporal, **kw):
    import ehgr_b200
    with contextlib.redirect_stdout(io.StringIO()):
        return ehgr_b200.TSN(83, 8, 'RGB', base_model='mobilenetv2', pretrain=None, dropout=0.5,
                             partial_bn=kw.pop("partial_bn", False), is_shift=(temporal != "none"), shift_div=8,
                             consensus_type='avg', fc_lr5=True, img_feature_dim=224,
                             temporal_module=("tsm" if temporal == "tsm" else "action"), **kw)


@pytest.mark.parametrize("temporal", ["none", "tsm", "action"])
def test_state_dict_keys_are_the_reference_checkpoint_contract(temporal):
    m = _tsn(temporal)
    sd = O.build_tsn_state(83, temporal, 8, seed=0)   # loads strict=True into the reference (make_golden.py)
    m.load_state_dict(sd, strict=True)
    assert sum(p.numel() for p in m.parameters()) == {"none": 2330195, "tsm": 2330195, "action": 2355455}[temporal]


def test_action_insertion_sites_and_structure():
    import ehgr_b200
    m = _tsn("action")
    sites = [i for i, f in enumerate(m.base_model.features)
             if isinstance(f, ehgr_b200.InvertedResidual) and isinstance(f.conv[0], ehgr_b200.Action)]
    assert sites == [3, 5, 6, 8, 9, 10, 12, 13, 15, 16]
    a = m.base_model.features[3].conv[0]
    assert (a.in_channels, a.out_channels, a.reduced_channels, a.fold, a.n_segment) == (24, 144, 1, 3, 8)
    assert len(m.base_model.features[1].conv) == 5 and len(m.base_model.features[2].conv) == 8
    assert m.base_model.last_layer_name == 'classifier' and isinstance(m.base_model.classifier, torch.nn.Dropout)
    t = _tsn("tsm")
    assert all(isinstance(t.base_model.features[i].conv[0], ehgr_b200.TemporalShift) for i in sites)
    assert t.base_model.features[3].conv[0].fold_div == 8


def test_optim_policies_group_sizes_match_reference():
    # MBv2+ACTION -> 1 / 0 / 51 / 0 / 104 / 80 / 20 / 1 / 1
    pol = _tsn("action").get_optim_policies()
    assert [len(g['params']) for g in pol] == [1, 0, 51, 0, 104, 80, 20, 1, 1]
    assert [g['lr_mult'] for g in pol] == [1, 2, 1, 2, 1, 1, 1, 5, 10]
    assert [g['decay_mult'] for g in pol] == [1, 0, 1, 0, 0, 1, 0, 1, 0]
    assert pol[4]['name'] == "BN scale/shift"


def test_partial_bn_freezes_all_but_first_bn():
    m = _tsn("action", partial_bn=True)
    with contextlib.redirect_stdout(io.StringIO()):
        m.train()
    bns = [b for b in m.base_model.modules() if isinstance(b, torch.nn.BatchNorm2d)]
    assert bns[0].training and not any(b.training for b in bns[1:])
    assert not bns[5].weight.requires_grad


def test_make_temporal_shift_rejects_unknown_backbones_like_the_reference():
    import ehgr_b200
    with contextlib.redirect_stdout(io.StringIO()):
        with pytest.raises(NotImplementedError):
            ehgr_b200.make_temporal_shift(torch.nn.Sequential(), 8)
        import torchvision
        r = torchvision.models.resnet18()
        ehgr_b200.make_temporal_shift(r, 8, n_div=8, place='blockres')
        assert isinstance(r.layer1[0].conv1, ehgr_b200.TemporalShift)


def test_no_cpu_fallback():
    import ehgr_b200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ehgr_b200.TemporalShift.shift(torch.zeros(8, 8, 2, 2), 8, fold_div=8)
    with pytest.raises(RuntimeError):   # the reference's .view error on a ragged segment count
        ehgr_b200.TemporalShift.shift(torch.zeros(7, 8, 2, 2), 4, fold_div=8)


def test_product_package_never_imports_the_oracle():
    from pathlib import Path
    import ehgr_b200
    pkg = Path(ehgr_b200.__file__).parent
    for f in pkg.rglob("*.py"):
        assert "oracle" not in f.read_text().replace("no oracle", ""), f


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle port on the host cores) prints one JSON line with the keys the
    driver reads; it must run without a GPU and without /root/reference."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    out = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-clips", "1"], capture_output=True, text=True, timeout=600, cwd=str(root))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["config"]["workload"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_normalize_u8_refuses_cpu_tensors_and_wrong_dtypes():
    """No CPU fallback anywhere in the product: the device-side input normalisation says so instead of computing."""
    import pytest
    import torch
    import ehgr_b200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ehgr_b200.train_step.normalize_u8(torch.zeros((1, 3, 4, 4), dtype=torch.uint8))


def test_mtmm_sd_wrapper_structure_and_policies():
    """models_MTMM_SD.TSN generalised to MobileNetV2: strict state_dict contract with the oracle's key set, ConvTranspose
    decoders as in models/models_MTMM_SD.py:226-249, policy groups (ConvTranspose2d counts as a convolution, :361)."""
    import contextlib, io
    import ehgr_b200 as E
    from oracle import ref_oracle as O
    with contextlib.redirect_stdout(io.StringIO()):
        m = E.tsn_mtmm_sd.TSN(83, 8, 'RGB', is_shift=True, partial_bn=False, base_model='mobilenetv2', shift_div=8, dropout=0.5,
                              img_feature_dim=224, pretrain=None, consensus_type='avg', fc_lr5=True, modal='rgb_depth',
                              temporal_module='tsm')
    m.load_state_dict(O.build_mtmm_sd_state(83, 'tsm', 8, 0), strict=True)
    assert [type(l).__name__ for l in m.global_decoder] == ['ConvTranspose2d', 'BatchNorm2d', 'ConvTranspose2d', 'BatchNorm2d',
                                                             'ConvTranspose2d', 'Sigmoid']
    assert [type(l).__name__ for l in m.local_decoder] == ['ConvTranspose2d', 'BatchNorm2d', 'ConvTranspose2d', 'Sigmoid']
    groups = {g['name']: g for g in m.get_optim_policies()}
    n_all = sum(len(g['params']) for g in groups.values())
    assert n_all == len(list(m.parameters()))
    assert len(groups['normal_bias']['params']) == 5            # the five ConvTranspose2d biases
    assert len(groups['lr5_weight']['params']) == 4 and len(groups['lr10_bias']['params']) == 4
    import pytest
    with pytest.raises(NotImplementedError):
        with contextlib.redirect_stdout(io.StringIO()):
            E.tsn_mtmm_sd.TSN(83, 8, 'RGB', base_model='mobilenetv2', pretrain=None, modal='rgb_depth_skeleton')


def test_depth_decoder_parser_recognises_the_reference_architecture_only():
    """fused.parse_depth_decoder maps models/models_MTMM.py:129-155 onto four conv3 stages (upsample flags on stages 2-4)
    and the 1x1 head; any other decoder (the ConvTranspose ones of models_MTMM_SD.py:226-249) is left to its modules."""
    import ehgr_b200 as E
    dec = E.tsn_mtmm.make_global_decoder(1280)
    stages, head = E.fused.parse_depth_decoder(dec)
    assert [(s.kind, s.conv.in_channels, s.conv.out_channels, s.up, s.relu6) for s in stages] == [
        ("conv3", 1280, 256, False, 2), ("conv3", 256, 64, True, 2), ("conv3", 64, 32, True, 2), ("conv3", 32, 32, True, 2)]
    assert head.in_channels == 32 and head.out_channels == 1 and head.bias is not None
    assert sorted(dec.state_dict()) == sorted(
        [f"{i}.weight" for i in (0, 4, 8, 12, 15)] + ["15.bias"] +
        [f"{i}.{k}" for i in (1, 5, 9, 13) for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")])
    assert E.fused.parse_depth_decoder(E.tsn_mtmm_sd.make_convt_decoder((1280, 256, 32, 1))) is None
    import torch.nn as nn
    assert E.fused.parse_depth_decoder(nn.Sequential(nn.Conv2d(8, 8, 3, padding=1, bias=False), nn.BatchNorm2d(8), nn.ReLU(),
                                                     nn.Upsample(scale_factor=2, mode="bilinear"), nn.Conv2d(8, 1, 1),
                                                     nn.Sigmoid())) is None
