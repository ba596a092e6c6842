"""CPU tests of the multi-GPU host logic with the gloo backend (world_size 2): row sharding covers the
table exactly once, owners agree with the ranges, and the data-parallel reductions (valid-target
count, loss, gradients summed over ranks) reproduce the single-process value on the global batch."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dr4sr_b200.dist import owner_of, shard_rows, split_batch, sum_over_ranks


def test_shard_rows_partition():
    for n, w in [(100_000, 8), (11_925, 2), (10_000_001, 8), (7, 8), (64, 4)]:
        ranges = shard_rows(n, w)
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [hi - lo for lo, hi in ranges]
        assert max(sizes) - min(sizes) <= 1
        ids = torch.arange(n) if n <= 200_000 else torch.randint(0, n, (200_000,))
        own = owner_of(ids, n, w)
        for r, (lo, hi) in enumerate(ranges):
            sel = ids[(ids >= lo) & (ids < hi)]
            assert bool((owner_of(sel, n, w) == r).all())
        assert int(own.min()) >= 0 and int(own.max()) < w


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import dr4sr_oracle as orc
        from dr4sr_b200.data.synthetic import synthetic_batch
        torch.manual_seed(0)
        torch.set_num_threads(1)
        N, D = 200, 64
        model = orc.OracleSASRec(N, embed_dim=D, dropout_rate=0.0).init_reference_style().train()
        full = synthetic_batch(8, 50, N, seed=5)
        # single-process reference on the global batch
        ref_loss = model.training_step(full)
        ref_loss.backward()
        ref_grads = [p.grad.clone() for p in model.parameters()]
        model.zero_grad()
        # data-parallel recipe used by BaseModel._step_forward/_step_backward: local sums, global n
        mine = split_batch(full, rank, world)
        q = model.forward(mine)
        pos, neg = orc.sampled_scores(q, model.item_embedding.weight, mine['item_id'], mine['neg_item'])
        valid = ~torch.isinf(pos)
        n = valid.sum().float().view(1)
        sum_over_ranks([n])
        local = (-(torch.nn.functional.logsigmoid(pos)[valid]).sum() + torch.nn.functional.softplus(neg[..., 0])[valid].sum()) / n
        local.backward()
        loss = local.detach().clone().view(1)
        grads = [p.grad for p in model.parameters()]
        sum_over_ranks([loss] + grads)
        ok = abs(float(loss) - float(ref_loss)) < 1e-6
        for g, r in zip(grads, ref_grads):
            ok = ok and bool(torch.allclose(g, r, rtol=1e-4, atol=1e-7))
        out[rank] = ok
    finally:
        dist.destroy_process_group()


def test_data_parallel_equals_single_process_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
