"""Order of gradient-done events and bucket launches of one SD / MTMM step (single GPU, collectives stubbed)."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ehgr_b200 as E
from ehgr_b200 import train_step as TS

workload = sys.argv[1] if len(sys.argv) > 1 else "sd"
dev = torch.device("cuda", 0)
torch.manual_seed(1)
common = dict(is_shift=True, partial_bn=False, base_model='mobilenetv2', shift_div=8, dropout=0.5, img_feature_dim=224,
              pretrain=None, consensus_type='avg', fc_lr5=True, temporal_module="tsm")
with contextlib.redirect_stdout(io.StringIO()):
    model = (E.tsn_sd.TSN(83, 8, 'RGB', **common) if workload == "sd" else E.tsn_mtmm.TSN(83, 8, 'RGB', modal='rgb_depth', **common))
model = model.to(dev).train()
cls = TS.SDTrainStep if workload == "sd" else TS.MTMMTrainStep
step = cls(model, compute_dtype=torch.bfloat16, use_graph=False)
gb = step.buckets
names = {id(p): n for n, p in model.named_parameters()}
log = []
gb.world = 2
gb._hooks = [p.register_post_accumulate_grad_hook(gb._on_grad) for p in gb.params]
orig_on = gb._on_grad
seen = set()
import traceback
def on_grad(p):
    if names[id(p)] in seen and len(seen) < 10**6 and not getattr(on_grad, "shown", False):
        on_grad.shown = True
        print("SECOND call for", names[id(p)], "grad is", None if p.grad is None else (p.grad.dtype, p.grad.data_ptr() == gb.flat.data_ptr() + 4 * gb._offset_of[id(p)]))
        traceback.print_stack(limit=8)
    seen.add(names[id(p)])
    log.append(("done", names[id(p)], gb._bucket_of[id(p)]))
    b = gb._bucket_of[id(p)]
    gb._left[b] -= 1
    if gb._left[b] == 0:
        log.append(("LAUNCH", b, None))
gb._on_grad = on_grad
for h in gb._hooks:
    h.remove()
gb._hooks = [p.register_post_accumulate_grad_hook(on_grad) for p in gb.params]
gb.finish = lambda: log.append(("finish", list(gb._left), None))
B = 2
rgb = torch.randn(B, 8, 3, 224, 224, device=dev)
depth = torch.rand(B, 8, 1, 224, 224, device=dev)
labels = torch.randint(0, 83, (B,), device=dev)
step.run(*((rgb, labels) if workload == "sd" else (rgb, depth, labels)))
torch.cuda.synchronize()
from collections import Counter
c = Counter(n for k, n, b in log if k == "done")
print("events", len(log), "params", len(gb.params), "double-counted:", [(n, k) for n, k in c.items() if k > 1][:20])
never = [names[id(p)] for p in gb.params if names[id(p)] not in c]
print("never done:", never[:30])
for e in log:
    if e[0] != "done":
        print(e)
idx = [i for i, e in enumerate(log) if e[0] == "LAUNCH"]
for i in idx:
    print("before launch", log[i], ":", [e[1] for e in log[max(0, i - 3):i]])
