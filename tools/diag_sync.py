"""Which parameters differ between data-parallel ranks after a few steps?  torchrun --nproc-per-node 2 tools/diag_sync.py"""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import ehgr_b200 as E

workload = sys.argv[1] if len(sys.argv) > 1 else "mtmm"
use_graph = (sys.argv[2] == "graph") if len(sys.argv) > 2 else False
torch.manual_seed(1)
common = dict(is_shift=True, partial_bn=False, base_model='mobilenetv2', shift_div=8, dropout=0.5, img_feature_dim=224,
              pretrain=None, consensus_type='avg', fc_lr5=True, temporal_module="tsm")
with contextlib.redirect_stdout(io.StringIO()):
    model = (E.tsn_sd.TSN(83, 8, 'RGB', **common) if workload == "sd" else E.tsn_mtmm.TSN(83, 8, 'RGB', modal='rgb_depth', **common))
model = model.to(dev).train()
cls = E.train_step.SDTrainStep if workload == "sd" else E.train_step.MTMMTrainStep
step = cls(model, compute_dtype=torch.bfloat16, use_graph=use_graph)
g = torch.Generator().manual_seed(100 + rank)
B = 4
for it in range(6):
    rgb = torch.randn(B, 8, 3, 224, 224, generator=g).to(dev)
    depth = torch.rand(B, 8, 1, 224, 224, generator=g).to(dev)
    labels = torch.randint(0, 83, (B,), generator=g).to(dev)
    loss = step.run(*((rgb, labels) if workload == "sd" else (rgb, depth, labels)))
    torch.cuda.synchronize()
    bad = []
    for name, p in model.named_parameters():
        mine = p.detach().float().flatten()
        other = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(other, mine)
        d = max(float((o - other[0]).abs().max()) for o in other)
        if d > 0:
            bad.append((name, d, float(other[0].abs().max())))
    # the reduced gradients themselves
    gbad = []
    for name, p in model.named_parameters():
        if p.grad is None:
            continue
        mine = p.grad.detach().float().flatten()
        other = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(other, mine)
        d = max(float((o - other[0]).abs().max()) for o in other)
        if d > 0:
            gbad.append((name, d))
    if rank == 0:
        print(f"step {it} loss {float(loss):.4f}: {len(bad)} parameters differ, {len(gbad)} reduced gradients differ", flush=True)
        for b in bad[:12]:
            print("   param", b)
        for b in gbad[:12]:
            print("   grad ", b)
        print("   in_sync:", None)
    ok = step.ranks_in_sync()
    if rank == 0:
        print("   ranks_in_sync():", ok, flush=True)
dist.barrier()
os._exit(0)
