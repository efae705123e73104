"""Data-parallel host logic on CPU: world_size-2 gloo run of the bucketed gradient all-reduce
(ehgr_b200.train_step.GradBuckets).  No GPU, no kernels: gradients are written by hand."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ehgr_b200
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(s)) for s in ((7, 3), (5,), (2, 2, 2), (11,), (1,))]
    gb = ehgr_b200.train_step.GradBuckets(params, n_buckets=3)
    assert gb.world == world and 2 <= len(gb.bounds) <= 3
    # slices are padded to 8 floats (32-byte boundaries: the fused chain writes weight gradients in place)
    assert sum(hi - lo for lo, hi in gb.bounds) == sum((p.numel() + 7) // 8 * 8 for p in params)
    assert all(lo % 8 == 0 for lo, _ in gb.bounds)
    for step in range(2):
        gb.zero()
        # a fake backward in reverse parameter order: every rank contributes (rank+1) * (index+1)
        for i, p in reversed(list(enumerate(params))):
            loss = (p * float((rank + 1) * (i + 1) + step)).sum()
            loss.backward()
        gb.finish()
        want = sum((r + 1) for r in range(world)) / world
        for i, p in enumerate(params):
            expect = ((want * (i + 1)) + step)
            assert torch.allclose(p.grad, torch.full_like(p, expect)), (rank, i, p.grad.flatten()[:3], expect)
            assert p.grad.data_ptr() >= gb.flat.data_ptr()          # grads are views of the flat buffer
    # gradient-sink protocol (fused.grad_sink): a producer writes into the flat slices itself and reports
    # finished parameters; autograd never sees those gradients
    gb.zero()
    for i, p in reversed(list(enumerate(params))):
        v = gb.view_for(p)
        assert v is not None and v.data_ptr() == p.grad.data_ptr() and float(v.abs().sum()) == 0.0
        v.add_(float((rank + 1) * (i + 1)))
        gb.mark_done([p])
    gb.finish()
    for i, p in enumerate(params):
        assert torch.allclose(p.grad, torch.full_like(p, want * (i + 1))), (rank, i)
    assert gb.view_for(torch.nn.Parameter(torch.zeros(3))) is None
    # a gradient detached by the caller (set_to_none / replaced tensor) is re-attached by the next zero()
    params[1].grad = None
    params[2].grad = torch.ones_like(params[2])
    gb.zero()
    assert params[1].grad is not None and params[1].grad.data_ptr() == gb.flat.data_ptr() + 4 * gb._offset_of[id(params[1])]
    assert params[2].grad.data_ptr() == gb.flat.data_ptr() + 4 * gb._offset_of[id(params[2])]
    assert float(params[2].grad.abs().sum()) == 0.0
    gb.finish()
    # unused parameter: its bucket is flushed by finish()
    gb.zero()
    (params[0] * 2.0).sum().backward()
    gb.finish()
    assert torch.allclose(params[0].grad, torch.full_like(params[0], 2.0))
    assert float(params[3].grad.abs().sum()) == 0.0
    # replicas that were seeded differently are brought to rank 0's state by the train step's constructor
    import contextlib, io
    torch.manual_seed(100 + rank)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ehgr_b200.tsn_mtmm.TSN(5, 8, 'RGB', base_model='mobilenetv2', pretrain=None, dropout=0.5, partial_bn=False,
                                   is_shift=True, fc_lr5=True, temporal_module='tsm', modal='rgb_depth')
    for b in m.buffers():
        if b.is_floating_point():
            b.add_(float(rank))
    w_before = m.new_fc.weight.detach().clone()
    gathered = [torch.empty_like(w_before) for _ in range(world)]
    dist.all_gather(gathered, w_before)
    assert not torch.equal(gathered[0], gathered[1])              # the seeds really differed
    step = ehgr_b200.train_step.MTMMTrainStep(m, compute_dtype=torch.float32)
    assert step.ranks_in_sync()
    assert torch.equal(m.new_fc.weight.detach(), gathered[0])
    rm = m.base_model.features[0][1].running_mean
    assert float(rm.abs().max()) == 0.0                            # rank 0's buffers (rank 1 had +1)
    rm.add_(float(rank))                                           # per-rank running statistics drift apart ...
    step.sync_bn_buffers()                                         # ... and are averaged before a checkpoint
    assert torch.allclose(rm, torch.full_like(rm, (world - 1) / 2))
    with torch.no_grad():
        if rank == 1:
            m.new_fc.weight.add_(1e-3)
    assert not step.ranks_in_sync()                                # the checksum notices a diverged replica
    if rank == 0:
        out.put("ok")
    dist.destroy_process_group()


def test_grad_buckets_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == "ok"


def test_bucket_assignment_with_one_dominant_parameter():
    """a parameter larger than a bucket's share must not make the assignment skip a bucket index."""
    import ehgr_b200
    for sizes in ((100, 100, 3_000_000, 100), (3_000_000,), (8, 8, 8, 8, 8, 8, 8, 8, 8), (5, 1_000_000, 5, 1_000_000, 5)):
        for nb in (1, 2, 3, 8):
            params = [torch.nn.Parameter(torch.zeros(n)) for n in sizes]
            gb = ehgr_b200.train_step.GradBuckets(params, n_buckets=nb)
            assert 1 <= len(gb.bounds) <= nb
            assert gb.bounds[0][0] == 0 and gb.bounds[-1][1] == gb.flat.numel()
            assert all(a[1] == b[0] for a, b in zip(gb.bounds, gb.bounds[1:]))       # contiguous, no gaps
            assert all(hi > lo for lo, hi in gb.bounds)
            assert sum(gb._need) == len(params)
            for p in params:
                lo, hi = gb.bounds[gb._bucket_of[id(p)]]
                assert lo <= gb._offset_of[id(p)] and gb._offset_of[id(p)] + p.numel() <= hi


def test_a_parameter_reports_done_once_per_step():
    """A sunk parameter reports through mark_done() AND (autograd runs its AccumulateGrad node afterwards) through the
    post-accumulate hook: the second report must not count, or a bucket mixing several autograd Functions launches
    its all-reduce before its last gradients exist (the 2-GPU SD divergence of round 2)."""
    import ehgr_b200
    params = [torch.nn.Parameter(torch.zeros(8)) for _ in range(6)]
    gb = ehgr_b200.train_step.GradBuckets(params, n_buckets=2)
    gb.world = 2                                   # pretend to be data-parallel; record launches instead of reducing
    launched = []
    gb._launch = lambda b: launched.append((b, sorted(id(p) for p in params if id(p) in gb._done)))
    order = list(reversed(params))                 # backward order
    b0 = [p for p in order if gb._bucket_of[id(p)] == 0]
    gb.zero()
    gb.mark_done(b0[:-1])                          # all but the last parameter of bucket 0
    for p in b0[:-1]:
        gb._on_grad(p)                             # ... whose hooks fire afterwards
    assert launched == []                          # bucket 0 still waits for its last parameter
    gb.mark_done(b0[-1:])
    assert [b for b, _ in launched] == [0]
    gb.zero()
    assert gb._done == set() and gb._left == gb._need


def test_sgd_groups_follow_reference_multipliers():
    import contextlib, io
    import ehgr_b200
    with contextlib.redirect_stdout(io.StringIO()):
        m = ehgr_b200.TSN(10, 8, 'RGB', base_model='mobilenetv2', pretrain=None, dropout=0.5, partial_bn=False,
                          is_shift=True, fc_lr5=True, temporal_module='tsm')
    opt = ehgr_b200.train_step.build_sgd(m, lr=0.01, momentum=0.9, weight_decay=5e-4)
    by_name = {g['name']: g for g in opt.param_groups}
    assert abs(by_name['lr5_weight']['lr'] - 0.05) < 1e-12 and abs(by_name['lr10_bias']['lr'] - 0.1) < 1e-12
    assert by_name['BN scale/shift']['weight_decay'] == 0 and abs(by_name['normal_weight']['weight_decay'] - 5e-4) < 1e-12
    ehgr_b200.train_step.adjust_learning_rate(0.01, opt, epoch=25, lr_steps=[20, 40])
    assert abs(by_name['normal_weight']['lr'] - 0.001) < 1e-12 and abs(by_name['lr5_weight']['lr'] - 0.005) < 1e-12
