"""Diagnostic: eager vs CUDA-graph train step trajectories (GPU)."""
import sys, io, contextlib
sys.path.insert(0, ".")
import torch
import ehgr_b200 as E
from oracle import ref_oracle as O


def make():
    sd0 = O.build_mtmm_state(83, "tsm", 8, seed=6)
    with contextlib.redirect_stdout(io.StringIO()):
        model = E.tsn_mtmm.TSN(83, 8, 'RGB', is_shift=True, partial_bn=False, base_model='mobilenetv2', shift_div=8,
                               dropout=0.5, img_feature_dim=224, pretrain=None, consensus_type='avg', fc_lr5=True,
                               modal='rgb_depth', temporal_module='tsm')
    model.load_state_dict(sd0, strict=True)
    model = model.cuda().train()
    for d in model.modules():
        if isinstance(d, torch.nn.Dropout):
            d.eval()
    return model


batches = [tuple(t.cuda() for t in O.synthetic_clip_batch(2, 8, 64, 83, seed=20 + i)) for i in range(5)]
traj = {}
for mode in ("eager", "eager2", "graph"):
    model = make()
    step = E.train_step.MTMMTrainStep(model, lr=1e-4, compute_dtype=torch.float32, use_graph=(mode == "graph"))
    rec = []
    for b in batches:
        loss = float(step.run(*b).item())
        sd = model.state_dict()
        rec.append((loss, {k: v.detach().clone() for k, v in sd.items() if v.is_floating_point()}))
    traj[mode] = rec
keys = ["base_model.features.0.0.weight", "base_model.features.1.conv.0.weight", "base_model.features.17.conv.6.weight", "new_fc.weight",
        "base_model.features.0.1.running_mean"]
for other in ("eager2", "graph"):
    print("==", other, "vs eager")
    for i in range(5):
        a, b = traj["eager"][i], traj[other][i]
        worst = max(((a[1][k] - b[1][k]).abs().max().item() / (a[1][k].abs().max().item() + 1e-12), k) for k in a[1])
        print(i, "loss", round(a[0], 5), round(b[0], 5), "worst", round(worst[0], 4), worst[1],
              [round(((a[1][k] - b[1][k]).abs().max() / (a[1][k].abs().max() + 1e-12)).item(), 4) for k in keys if k in a[1]])
d = traj["eager"]
print("eager step-to-step change", [round(((d[i + 1][1][keys[0]] - d[i][1][keys[0]]).abs().max() / d[i][1][keys[0]].abs().max()).item(), 4) for i in range(4)])
