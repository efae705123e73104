"""Diagnostic (GPU): per-unit error of the fused chain vs the fp64 oracle, fp32 / bf16-SIMT / bf16-tcgen05."""
import contextlib, io, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ehgr_b200 as E
from oracle import ref_oracle as O

sd0 = O.build_mtmm_state(83, "tsm", 8, seed=2)
with contextlib.redirect_stdout(io.StringIO()):
    model = E.tsn_mtmm.TSN(83, 8, 'RGB', is_shift=True, partial_bn=False, base_model='mobilenetv2', shift_div=8, dropout=0.5,
                           img_feature_dim=224, pretrain=None, consensus_type='avg', fc_lr5=True, modal='rgb_depth', temporal_module='tsm')
model.load_state_dict(sd0, strict=True)
model = model.cuda().train()
rgb, depth, labels = O.synthetic_clip_batch(1, 8, 224, 83, seed=4)
sd = O.clone_state(sd0, dtype=torch.float64, requires_grad=False)
taps = {i: None for i in range(1, 19)}
with torch.no_grad():
    O.tsn_forward(rgb.double(), sd, 8, "tsm", 8, True, taps=taps)
    dref = O.global_decoder(taps[18], sd, True)
x = rgb.view(-1, 3, 224, 224).cuda()
for name, dt, eng in (("fp32", torch.float32, 0), ("bf16-simt", torch.bfloat16, 1), ("bf16-tc", torch.bfloat16, 2)):
    model.load_state_dict(sd0, strict=True)
    with torch.no_grad(), E.fused.compute_dtype(dt), E.fused.gemm_engine(eng):
        outs = E.fused.mobilenet_v2_features(model.base_model, x, taps=list(range(1, 19)))
        errs = []
        for i, o in zip(range(1, 19), outs):
            r = taps[i]
            errs.append(((o.double().cpu() - r).abs().max() / r.abs().max()).item())
        d = model.global_decoder(outs[-1])
        dd = (d.double().cpu() - dref).abs().max().item()
        d2 = model.global_decoder(taps[18].float().cuda())
        dd2 = (d2.double().cpu() - dref).abs().max().item()
    print(name, " ".join(f"{e:.1e}" for e in errs), "| decoder abs err", f"{dd:.2e}", "| decoder(oracle f18)", f"{dd2:.2e}")

# how does plain PyTorch bf16 autocast do on the same case? (oracle functions on the GPU under autocast)
def rms(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()

sdg = {k: v.float().cuda() for k, v in O.clone_state(sd0, requires_grad=False).items()}
tg = {i: None for i in range(1, 19)}
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    O.tsn_forward(rgb.cuda(), sdg, 8, "tsm", 8, True, taps=tg)
    dg = O.global_decoder(tg[18], sdg, True)
print("torch-autocast-bf16 max:", " ".join(f"{((tg[i].double().cpu()-taps[i]).abs().max()/taps[i].abs().max()).item():.1e}" for i in range(1, 19)),
      "| decoder abs", f"{(dg.double().cpu()-dref).abs().max().item():.2e}")
print("torch-autocast-bf16 rms:", " ".join(f"{rms(tg[i], taps[i]):.1e}" for i in range(1, 19)))
model.load_state_dict(sd0, strict=True)
with torch.no_grad(), E.fused.compute_dtype(torch.bfloat16), E.fused.gemm_engine(2):
    outs = E.fused.mobilenet_v2_features(model.base_model, x, taps=list(range(1, 19)))
print("ehgr-bf16-tc        rms:", " ".join(f"{rms(o, taps[i]):.1e}" for i, o in zip(range(1, 19), outs)))
