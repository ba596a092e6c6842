"""Multi-GPU parity (needs >= 2 GPUs; run with `gpurun --gpus 2 -- pytest tests/test_cuda_sharded.py -m gpu`):
the row-sharded table + data-parallel encoder on 2 ranks x B sequences computes the same loss, gradients,
Adam update and top-k as one process on the 2B global batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _worker(rank, world, port, model_name, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        from dr4sr_b200.data.synthetic import synthetic_batch
        from dr4sr_b200.dist import split_batch
        from dr4sr_b200.utils.config import SyntheticCatalog, default_config
        if model_name == 'SASRec':
            from dr4sr_b200.model.sasrec import SASRec as Model
        else:
            from dr4sr_b200.model.gru4rec import GRU4Rec as Model
        N, D, B = 5003, 128, 24
        errs = {}

        def build(shard):
            cfg = default_config(model_name, model__embed_dim=D, model__dropout_rate=0.0, train__device=str(dev),
                                 train__weight_decay=0.0)
            if shard:
                cfg['train']['table_shard'] = (rank, world)
            torch.manual_seed(7)
            m = Model(cfg, [SyntheticCatalog(N)] * 3)
            m._init_model()
            return m

        ref = build(False).train()                       # whole table, whole batch, on this rank's GPU
        full = {k: v.to(dev) for k, v in synthetic_batch(world * B, 50, N, seed=3).items()}
        sh = build(True)
        lo, hi = sh._shard_rows
        sh.item_embedding.weight.data.copy_(ref.item_embedding.weight.data[lo:hi])
        sh._flat.copy_(ref._flat)
        sh.enable_sharded_table(dist.group.WORLD)
        sh.train()
        mine = split_batch(full, rank, world)

        ref.optimizer.zero_grad()
        lref = ref.training_step(full)
        lref.backward()
        sh.optimizer.zero_grad()
        lsh = sh.training_step(mine)
        lsh.backward()
        errs['loss'] = abs(float(lsh.detach()) - float(lref.detach())) / abs(float(lref.detach()))
        errs['flat_grad'] = _rel(sh._flat_grad, ref._flat_grad)
        errs['table_grad'] = _rel(sh.item_embedding.weight.grad, ref.item_embedding.weight.grad[lo:hi])
        ref.optimizer.step()
        sh.optimizer.step()
        errs['table_after_adam'] = float((sh.item_embedding.weight.data - ref.item_embedding.weight.data[lo:hi]).abs().max())
        errs['flat_after_adam'] = float((sh._flat - ref._flat).abs().max())

        ref.eval(); sh.eval()
        sh.item_embedding.weight.data.copy_(ref.item_embedding.weight.data[lo:hi])   # identical parameters for the id check
        sh._flat.copy_(ref._flat)
        ev = {k: v.to(dev) for k, v in synthetic_batch(world * B, 50, N, seed=4, eval_mode=True, with_neg=False).items()}
        s_ref, i_ref = ref.topk(ev, 100, ev['user_hist'])
        ev_mine = split_batch(ev, rank, world)
        s_sh, i_sh = sh.topk(ev_mine, 100, ev_mine['user_hist'])
        errs['topk_scores'] = _rel(s_sh, s_ref[rank::world])
        errs['topk_ids_equal'] = float((i_sh == i_ref[rank::world]).float().mean())
        out[rank] = errs
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('model_name', ['SASRec', 'GRU4Rec'])
def test_sharded_table_equals_single_process(model_name):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, model_name, out), nprocs=world, join=True)
    res = dict(out)
    assert set(res) == {0, 1}
    for r, e in res.items():
        assert e['loss'] < 1e-5, (r, e)
        assert e['flat_grad'] < 1e-3, (r, e)
        assert e['table_grad'] < 1e-3, (r, e)
        assert e['table_after_adam'] < 1e-4 and e['flat_after_adam'] < 1e-4, (r, e)
        assert e['topk_scores'] < 1e-5, (r, e)
        assert e['topk_ids_equal'] > 0.999, (r, e)
