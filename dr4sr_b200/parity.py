"""Self-check of the row-sharded multi-GPU path against the single-process CUDA path, callable from any launched rank
set (tests/test_cuda_sharded.py spawns 2 ranks; bench.py runs it on the driver's N ranks before timing, so the sharded
layout is exercised on the same box that produces the scaling numbers).

The single-process side is the plain CUDA model (itself pinned to the oracle and the reference goldens by the `-m gpu`
tests); nothing here touches `oracle/`.  Every rank builds (a) the full table + the global batch of world x B sequences
and (b) its shard + its rows of the batch, and compares loss, encoder gradients, owned table-gradient rows, parameters
after one Adam step, and full-catalog top-k (ids must be EQUAL: same parameters, same kernels, same per-row arithmetic).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.distributed as dist


def _rel(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def sharded_parity_check(group, dev: torch.device, model_name: str = 'SASRec', N: int = 5003, D: int = 128, B: int = 24,
                         k: int = 100, mode: str = 'a2a') -> Dict[str, float]:
    """-> {'loss_rel', 'grad_rel', 'table_grad_rel', 'table_after_adam_abs', 'flat_after_adam_abs', 'topk_scores_rel',
    'topk_ids_equal', 'step_freed_by_refcount'} maximised (minimised for the last two) over the ranks of `group`."""
    from .data.synthetic import synthetic_batch
    from .dist import split_batch
    from .utils.config import SyntheticCatalog, default_config
    if model_name == 'SASRec':
        from .model.sasrec import SASRec as Model
    elif model_name == 'GRU4Rec':
        from .model.gru4rec import GRU4Rec as Model
    else:
        raise ValueError(model_name)
    rank, world = dist.get_rank(group), dist.get_world_size(group)

    def build(shard: bool):
        cfg = default_config(model_name, model__embed_dim=D, model__dropout_rate=0.0, train__device=str(dev), train__weight_decay=0.0)
        if shard:
            cfg['train']['table_shard'] = (rank, world)
        torch.manual_seed(7)
        m = Model(cfg, [SyntheticCatalog(N)] * 3)
        m._init_model()
        return m

    ref = build(False).train()                       # whole table, whole batch, on this rank's GPU
    full = {k_: v.to(dev) for k_, v in synthetic_batch(world * B, 50, N, seed=3).items()}
    sh = build(True)
    lo, hi = sh._shard_rows
    sh.item_embedding.weight.data.copy_(ref.item_embedding.weight.data[lo:hi])
    if mode == 'peer':                               # rows read / gradient rows added directly in the owner's HBM (peer memory)
        sh.enable_peer_table(group)
    else:                                            # rows exchanged by all-to-all
        sh.enable_sharded_table(group)               # (both broadcast rank 0's encoder: identical on every rank, same seed)
    sh._flat.copy_(ref._flat)
    sh.train()
    sh.config['train']['early_loss_read'] = True     # loss_value(): the global loss right after the forward
    mine = split_batch(full, rank, world)

    errs: Dict[str, float] = {}
    ref.optimizer.zero_grad()
    lref = ref.training_step(full)
    lref.backward()
    sh.optimizer.zero_grad()
    lsh = sh.training_step(mine)
    lsh.backward()
    errs['loss_rel'] = abs(float(lsh.detach()) - float(lref.detach())) / abs(float(lref.detach()))
    errs['loss_value_rel'] = abs(sh.loss_value() - float(lref.detach())) / abs(float(lref.detach()))   # early side-stream read, global
    errs['grad_rel'] = _rel(sh._flat_grad, ref._flat_grad)
    errs['table_grad_rel'] = _rel(sh.item_embedding.weight.grad, ref.item_embedding.weight.grad[lo:hi])
    ref.optimizer.step()
    sh.optimizer.step()
    # the step's tensors must die by reference count (no cycle through the autograd node): a cycle leaves them to the cyclic
    # GC, the allocator pool keeps growing, and each cudaMalloc under peer mappings stalls the host for milliseconds
    import gc
    import weakref
    was = gc.isenabled()
    gc.disable()
    w = weakref.ref(lsh)
    del lsh
    errs['step_freed_by_refcount'] = 1.0 if w() is None else 0.0
    if was:
        gc.enable()
    errs['table_after_adam_abs'] = float((sh.item_embedding.weight.data - ref.item_embedding.weight.data[lo:hi]).abs().max())
    errs['flat_after_adam_abs'] = float((sh._flat - ref._flat).abs().max())

    ref.eval(); sh.eval()
    # one set of parameters for everybody: the ranks' single-process models took their Adam step independently (float atomics
    # in the scatter-add: equal to round-off only), and the sharded table is assembled from every rank's slice
    dist.broadcast(ref.item_embedding.weight.data, src=dist.get_global_rank(group, 0), group=group)
    dist.broadcast(ref._flat, src=dist.get_global_rank(group, 0), group=group)
    sh.item_embedding.weight.data.copy_(ref.item_embedding.weight.data[lo:hi])   # identical parameters for the id check
    sh._flat.copy_(ref._flat)
    ev = {k_: v.to(dev) for k_, v in synthetic_batch(world * B, 50, N, seed=4, eval_mode=True, with_neg=False).items()}
    # the SAME queries through both paths (this rank's rows of the batch): the encoder's arithmetic for a row depends on where
    # the row sits in its 128-row tile (tensor-core accumulation order), so only equal batches give bit-equal query vectors
    ev_mine = split_batch(ev, rank, world)
    s_ref, i_ref = ref.topk(ev_mine, k, ev_mine['user_hist'])
    s_sh, i_sh = sh.topk(ev_mine, k, ev_mine['user_hist'])
    errs['topk_scores_rel'] = _rel(s_sh, s_ref)
    errs['topk_ids_equal'] = float((i_sh == i_ref).float().mean())

    # worst over ranks
    keys = sorted(errs)
    mins = ('topk_ids_equal', 'step_freed_by_refcount')         # reported as the minimum over ranks
    t = torch.tensor([errs[k_] if k_ not in mins else -errs[k_] for k_ in keys], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    out = {k_: (float(v) if k_ not in mins else -float(v)) for k_, v in zip(keys, t.tolist())}
    out.update(model=model_name, world=world, num_items=N, batch_per_rank=B, mode=mode)
    del ref, sh
    torch.cuda.empty_cache()
    return out
