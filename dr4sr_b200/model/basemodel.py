"""BaseModel: the trainer-is-the-model surface of the reference (reference model/basemodel.py:19-407),
re-hosted on libdr4sr.  Same constructor, hooks and batch contract:

    Model(config, dataset_list); fit(); evaluate(); _init_model(); _neg_sampling(batch);
    training_step(batch, reduce=True, return_query=False); forward(batch); topk(batch, k, user_h);
    set_eval_domain(domain); load_checkpoint(path)

What differs is underneath: negatives come from a counter-based kernel instead of a B x N multinomial
(basemodel.py:50-61), scoring + BCE + their backward are one kernel (basemodel.py:205-210,
loss_func.py:9-35), the table gradient is one scatter-add instead of three dense [N, D] buffers, Adam
is a single fused pass, and top-k never copies the domain item list to the device per batch
(basemodel.py:358-359).
"""
from __future__ import annotations

import logging
import os
import time
from collections import defaultdict
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .. import engine as _engine
from ..optim import FlatGroup, FusedAdam
from .loss_func import BinaryCrossEntropyLoss, BPRLoss


def normal_initialization(module: nn.Module, initial_range: float = 0.02) -> None:
    """Reference utils/utils.py:70-81."""
    if isinstance(module, nn.Embedding):
        module.weight.data.normal_(mean=0.0, std=initial_range)
        if module.padding_idx is not None:
            module.weight.data[module.padding_idx].zero_()
    elif isinstance(module, nn.Linear):
        module.weight.data.normal_(mean=0.0, std=initial_range)
        if module.bias is not None:
            module.bias.data.zero_()
    elif isinstance(module, nn.LayerNorm):
        module.bias.data.zero_()
        module.weight.data.fill_(1.0)


class _TrainStep(torch.autograd.Function):
    """training_step as one autograd node.  forward = the model's forward + loss kernels, backward = its
    backward kernels; gradients are written by the kernels into the model's flat buffers and published as
    ``.grad`` -- nothing flows back through autograd edges."""

    @staticmethod
    def forward(ctx, table, model, batch, reduce, return_query):
        model._check_flat()
        loss, query, state = model._step_forward(batch, reduce, return_query)
        eng = model.engine
        eng.fwd_token += 1
        ctx.model, ctx.state, ctx.reduce, ctx.token = model, state, reduce, eng.fwd_token
        ctx.set_materialize_grads(False)
        return (loss, query) if return_query else loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dloss, dquery=None):
        model = ctx.model
        if ctx.token != model.engine.fwd_token:
            raise RuntimeError('dr4sr_b200: backward() of a stale training_step (the engine keeps the activations of the '
                               'most recent forward only)')
        if dloss is None:
            raise RuntimeError('dr4sr_b200: training_step loss received no gradient')
        opt = getattr(model, 'optimizer', None)
        if hasattr(opt, 'pending_backwards'):
            if opt.pending_backwards >= 1:
                raise RuntimeError('dr4sr_b200: a second backward() without optimizer.step() or optimizer.zero_grad() in between: the '
                                   'kernels overwrite the gradient buffers, they do not accumulate across backward() calls '
                                   '(torch would); step or zero_grad first')
            opt.pending_backwards += 1
        model._step_backward(ctx.state, ctx.reduce, dloss.contiguous(), dquery)
        model._publish_grads()
        return None, None, None, None, None


class BaseModel(nn.Module):
    def __init__(self, config: Dict, dataset_list: List) -> None:
        super().__init__()
        self.config = config
        self.ckpt_path = None
        self.logger = logging.getLogger('CDR')
        self.dataset_list = dataset_list
        self.device = config['train']['device']
        self.fuid, self.fiid = 'user_id', 'item_id'
        head = dataset_list[0]
        self.domain_name_list = head.domain_name_list
        self.domain_user_mapping = head.domain_user_mapping
        self.domain_item_mapping = head.domain_item_mapping
        self.training_time = 0.0
        self.inference_time = 0.0
        self.embed_dim = config['model']['embed_dim']
        self.max_seq_len = config['data']['max_seq_len']
        self.num_users = head.num_users
        self.num_items = head.num_items
        # config['train']['table_shard'] = (rank, world): this process holds rows [lo, hi) of the item table only
        # (multi-GPU, see dr4sr_b200/sharded.py); absent => the whole table, as in the reference (basemodel.py:42)
        self._shard = None
        self._peer = None
        self._shard_rows = None
        ts = config['train'].get('table_shard')
        if ts is not None:
            from ..dist import shard_rows
            self._shard_rows = shard_rows(self.num_items, int(ts[1]))[int(ts[0])]
            lo, hi = self._shard_rows
            rows, pad = hi - lo, (0 if lo == 0 else None)
        else:
            rows, pad = self.num_items, 0
        # config['train']['table_init_device'] (optional): allocate + initialise the table there directly (a 10M-row table
        # costs ~10 s of host RNG otherwise); absent => on the host like the reference, then moved by _init_model
        self.item_embedding = nn.Embedding(rows, self.embed_dim, padding_idx=pad, device=config['train'].get('table_init_device'))
        self.eval_domain = self.domain_name_list[0]
        self.engine = None
        self._dead_cache: Dict[str, torch.Tensor] = {}
        self._neg_step = 0

    @staticmethod
    def _get_dataset_class(config):
        """Reference model/basemodel.py:63-77; the shipped configs all use 'general' = SeparateDataset."""
        from ..data.dataset import SeparateDataset
        if config['data']['dataset_class'] == 'general':
            return SeparateDataset
        raise NotImplementedError(f"dataset_class {config['data']['dataset_class']!r}: only 'general' is used by the shipped configs")

    # ---- hooks a subclass provides ----------------------------------------------------------
    def _build_engine(self) -> None:
        raise NotImplementedError

    def _flat_parameters(self) -> List[nn.Parameter]:
        """Non-table parameters in the order of the engine's flat layout."""
        raise NotImplementedError

    def forward(self, batch):
        raise NotImplementedError

    # ---- setup ------------------------------------------------------------------------------
    def _init_model(self, train_data=None):
        self.apply(normal_initialization)
        self.to(self.device)
        if torch.device(self.device).type != 'cuda':
            raise _engine._lib.Dr4srError("dr4sr_b200 runs on CUDA only (config['train']['device'] must be a GPU)")
        self._build_engine()
        self._flatten()
        self.optimizer = self._get_optimizers()
        self.loss_fn = self._get_loss_func()
        self.engine.loss_kind = {'bce': 0, 'bpr': 1}[self.loss_fn.kind]      # DR4SR_LOSS_* of include/dr4sr.h
        self.engine.deterministic_scatter = bool(self.config['train'].get('deterministic_scatter', False))

    def _flatten(self) -> None:
        """Re-home the encoder parameters as views of one flat fp32 buffer (the C ABI's layout) and
        allocate the gradient buffers.  state_dict keys and shapes are unchanged."""
        params = self._flat_parameters()
        n = sum(p.numel() for p in params)
        if n != self.engine.param_count:
            raise _engine._lib.Dr4srError(f'flat layout mismatch: python {n} vs C {self.engine.param_count}')
        dev = self.item_embedding.weight.device
        self._flat = torch.empty(n, dtype=torch.float32, device=dev)
        # every gradient lives in ONE buffer [table gradient | encoder gradient | loss slot], so that a data-parallel step
        # needs a single all-reduce (replicated table) or one over the tail (row-sharded table)
        tn = self.item_embedding.weight.numel()
        self._comm = torch.zeros(tn + n + 1, dtype=torch.float32, device=dev)
        self._flat_grad = self._comm[tn:tn + n]
        self._grad_views = []
        off = 0
        for p in params:
            k = p.numel()
            self._flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self._flat[off:off + k].view(p.shape)
            self._grad_views.append(self._flat_grad[off:off + k].view(p.shape))
            off += k
        self._flat_params = params
        self._table_grad = self._comm[:tn].view_as(self.item_embedding.weight.data)
        old_table = getattr(self, '_table_group', None)
        self._table_group: Optional[FlatGroup] = None
        opt = getattr(self, 'optimizer', None)
        if isinstance(opt, FusedAdam):
            # parameters were moved / re-created after _init_model: re-point the optimizer's flat groups at the new
            # buffers and carry the Adam moments over, so the live parameters keep being the ones that are updated
            for g in opt.flat_groups:
                new_p, new_g = (self.item_embedding.weight.data, self._table_grad) if g is old_table else (self._flat, self._flat_grad)
                if g.m.shape != new_p.shape:
                    raise _engine._lib.Dr4srError(f'parameter buffer {g.name!r} changed shape after _init_model; rebuild the optimizer')
                g.param, g.grad = new_p, new_g
                g.m, g.v = g.m.to(new_p.device), g.v.to(new_p.device)
                g.dirty = False
            self._table_group = old_table

    def _check_flat(self) -> None:
        p0 = self._flat_params[0]
        if p0.data_ptr() != self._flat.data_ptr() or self._table_grad.device != self.item_embedding.weight.device:
            self._flatten()          # parameters were moved / re-created after _init_model

    def _get_optimizers(self):
        t = self.config['train']
        name, lr, wd = t['optimizer'].lower(), t['learning_rate'], t['weight_decay']
        if name == 'adam':
            self._table_group = FlatGroup(self.item_embedding.weight.data, self._table_grad, 'table', zero_grad_in_step=True)
            groups = [FlatGroup(self._flat, self._flat_grad, 'encoder'), self._table_group]
            return FusedAdam(self.parameters(), groups, lr=lr, weight_decay=wd)
        if name == 'sgd':
            return torch.optim.SGD(self.parameters(), lr=lr, weight_decay=wd)
        if name == 'adagrad':
            return torch.optim.Adagrad(self.parameters(), lr=lr, weight_decay=wd)
        if name == 'rmsprop':
            return torch.optim.RMSprop(self.parameters(), lr=lr, weight_decay=wd)
        raise NotImplementedError(f'optimizer {name!r}')

    def _get_loss_func(self):
        kind = self.config['model']['loss_fn']
        if kind == 'bce':
            return BinaryCrossEntropyLoss()
        if kind == 'bpr':
            return BPRLoss()
        raise NotImplementedError(kind)

    # ---- data parallel (no reference counterpart: the reference is single-device) -----------------
    def enable_data_parallel(self, group) -> None:
        """Replicated parameters, batch sharded over the ranks of `group`.  The loss normaliser n and the
        gradients are summed over ranks, so N ranks x B sequences compute exactly the single-process step
        on the N*B global batch (up to fp32 summation order)."""
        self._dp_group = group
        self._sync_replicas(group, table=True)

    def _sync_replicas(self, group, table: bool) -> None:
        """Every rank starts from rank 0's replicated parameters (the callers need not seed the ranks identically)."""
        import torch.distributed as dist
        self._check_flat()
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast(self._flat, src=src, group=group)
        if table:
            dist.broadcast(self.item_embedding.weight.data, src=src, group=group)
        extra = [p.data for p in self.parameters() if p is not self.item_embedding.weight and all(p is not q for q in self._flat_params)]
        for t in extra:                                   # parameters outside the flat encoder buffer (none for the shipped models)
            dist.broadcast(t, src=src, group=group)

    def enable_sharded_table(self, group) -> None:
        """Row-sharded item table (config['train']['table_shard'] must have sized the embedding as the shard):
        rows are fetched / their gradients returned by all-to-all, Adam runs on the shard; the encoder stays
        data parallel."""
        from ..sharded import ShardedTable
        if self._shard_rows is None:
            raise _engine._lib.Dr4srError("enable_sharded_table needs config['train']['table_shard'] = (rank, world)")
        self._dp_group = group
        self._shard = ShardedTable(self.num_items, self.embed_dim, group, self.item_embedding.weight.device)
        assert (self._shard.lo, self._shard.hi) == tuple(self._shard_rows)
        self._sync_replicas(group, table=False)           # the encoder is replicated; each rank keeps its own table rows

    def enable_peer_table(self, group) -> None:
        """Row-sharded item table over peer memory (config['train']['table_shard'] must have sized the embedding as this
        rank's shard): kernels read rows / add gradient rows directly in the owner's HBM over NVLink, Adam runs on the shard;
        the encoder stays data parallel.  See dr4sr_b200/peer.py.  One backward per optimizer step."""
        from ..peer import PeerTable
        from ..sharded import ShardedTable
        if self._shard_rows is None:
            raise _engine._lib.Dr4srError("enable_peer_table needs config['train']['table_shard'] = (rank, world)")
        if getattr(self.engine, '_name', '') != 'dr4sr_sasrec':
            raise _engine._lib.Dr4srError(f'the peer-memory table is implemented for SASRec; use enable_sharded_table (all-to-all) for '
                                          f'{type(self).__name__}')
        if self.engine.deterministic_scatter:
            raise _engine._lib.Dr4srError('deterministic_scatter is not available with the peer-memory table (gradient rows of several '
                                          'ranks meet in the owner\'s HBM through float atomics); use the all-to-all or replicated layout')
        self._dp_group = group
        self._sync_replicas(group, table=False)
        self._table_grad.zero_()
        if self._table_group is not None:
            self._table_group.dirty = False
        n = self._flat_grad.numel()
        self._peer = PeerTable(self.num_items, self.embed_dim, group, self.item_embedding.weight.data, self._table_grad, n_stage=n + 1,
                               collectives=self.config['train'].get('peer_collectives', 'peer'))
        self._peer_eval = ShardedTable(self.num_items, self.embed_dim, group, self.item_embedding.weight.device)   # top-k merge only

    def _grad_out(self) -> torch.Tensor:
        """Where the backward writes the flat encoder gradient: the optimizer's buffer, or -- peer collectives -- the staging
        buffer the other ranks read (the all-reduce then lands in the optimizer's buffer)."""
        if self._peer is not None and self._peer.coll:
            return self._peer.stage[: self._flat_grad.numel()]
        return self._flat_grad

    def _rows_for(self, bufs, in_ids, item_id, neg):
        """(table, in_ids, item_id, neg) the kernels run on: the parameter itself, the peer-sharded table, or the staged
        local rows of the all-to-all variant."""
        if self._peer is not None:
            self._peer.check_unmoved(self.item_embedding.weight.data, self._table_grad)
            return self._peer, in_ids, item_id, neg
        if self._shard is None:
            return self.item_embedding.weight.data, in_ids, item_id, neg
        return self._shard.fetch(self.item_embedding.weight.data, bufs, in_ids, item_id, neg)

    def _scatter_target(self):
        if self._peer is not None:            # cleared by the Adam pass only: other ranks add to this shard at any time of the step
            if self._table_group is not None:
                self._table_group.dirty = True
            return self._peer
        return self._table_grad_buffer() if self._shard is None else self._shard.local_grad()

    def _finish_table_grad(self, tg: torch.Tensor) -> None:
        if self._shard is None:
            self._dp_sum(tg)
        else:
            self._shard.push_grads(self._table_grad_buffer())

    def _dp_count_async(self, counts_slot: torch.Tensor):
        """Global number of valid targets: summed over ranks on NCCL's stream while the encoder forward runs; the
        returned handle's wait() orders the scoring kernel behind it (None when not data parallel)."""
        grp = getattr(self, '_dp_group', None)
        if grp is None:
            return None
        import torch.distributed as dist
        return dist.all_reduce(counts_slot, op=dist.ReduceOp.SUM, group=grp, async_op=True)

    def _reduce_grads(self, tg: torch.Tensor, loss: Optional[torch.Tensor]) -> None:
        """End of a data-parallel backward: ONE all-reduce over [table gradient | encoder gradient | loss] (replicated
        table), or the row push of the sharded table plus one all-reduce over [encoder gradient | loss].  `loss` (the
        forward's local partial sum, divided by the global n) becomes the global loss in place."""
        grp = getattr(self, '_dp_group', None)
        if grp is None:
            return
        import torch.distributed as dist
        tn = self._table_grad.numel()
        if self._peer is not None and self._peer.coll:
            # gradient rows were added in their owners' HBM by the scatter kernels; the encoder gradients (and the loss) sit in the
            # staging buffer: barrier ("every rank's backward and scatter are done") + sum over the ranks' buffers, no NCCL
            n = self._flat_grad.numel()
            if loss is not None:
                self._peer.stage[n:n + 1].copy_(loss.detach().view(1))
            self._peer.allreduce_stage(n + 1, self._comm[tn:])
        elif self._peer is not None:
            if loss is not None:
                self._comm[-1:].copy_(loss.detach().view(1))
            # the rows were added in their owners' HBM by the scatter kernel; this all-reduce is also the barrier
            # "every rank's scatter has finished" that the Adam pass on the shard needs
            dist.all_reduce(self._comm[tn:], op=dist.ReduceOp.SUM, group=grp)
        elif self._shard is None and tg is self._table_grad:
            if loss is not None:
                self._comm[-1:].copy_(loss.detach().view(1))
            dist.all_reduce(self._comm, op=dist.ReduceOp.SUM, group=grp)
        else:
            if loss is not None:
                self._comm[-1:].copy_(loss.detach().view(1))
            self._finish_table_grad(tg)
            dist.all_reduce(self._comm[tn:], op=dist.ReduceOp.SUM, group=grp)
        if loss is not None:
            loss.detach().view(1).copy_(self._comm[-1:])
            if getattr(self, '_loss_late', False):
                self._loss_evt.record()
                self._loss_late = False

    # ---- host read of the loss without draining the stream ---------------------------------------
    def _mark_loss(self, loss: torch.Tensor, final: bool = True) -> None:
        """Called by the step's forward once the kernels that produce `loss` are enqueued.  `final=False`: data parallel with
        the rank-local partial (it becomes global inside the backward's gradient all-reduce): the global value for
        `loss_value()` is formed right away by a 1-float all-reduce on the read stream -- beside the backward, never waited
        for by the compute stream."""
        if getattr(self, '_loss_evt', None) is None:
            dev = loss.device
            self._loss_evt = torch.cuda.Event()
            self._read_stream = torch.cuda.Stream(device=dev)
            self._loss_pin = torch.zeros(1, dtype=torch.float32).pin_memory()
            self._loss_glob = torch.zeros(1, dtype=torch.float32, device=dev)
        self._loss_ref = loss.detach().view(1)
        self._loss_evt.record()
        self._loss_late = False
        if not final and self.config['train'].get('early_loss_read', False):
            # (a collective: the switch is configuration, identical on every rank -- never a per-rank decision)
            import torch.distributed as dist
            s = self._read_stream
            s.wait_event(self._loss_evt)
            with torch.cuda.stream(s):
                self._loss_glob.copy_(self._loss_ref)
                dist.all_reduce(self._loss_glob, op=dist.ReduceOp.SUM, group=self._dp_group)
            self._loss_ref = self._loss_glob
        elif not final:
            self._loss_late = True               # _reduce_grads records the event again once the value is global

    def loss_value(self) -> float:
        """Host value of the most recent `training_step(reduce=True)` loss (the global one under data parallelism), read as
        soon as the kernels that produce it have finished: an event-gated copy on a side stream.  The backward and the
        optimizer step of the same batch (already enqueued) keep running and the host goes on enqueuing the next batch --
        `float(loss)` would wait for all of it.  (The reference's trainer never reads the loss inside the epoch,
        model/basemodel.py:199,217-224.)  Data parallel: with config['train']['early_loss_read'] the global value is formed by a
        1-float all-reduce on the read stream right after the forward; without it the value is ready when the backward's
        gradient all-reduce is (call after `backward()`)."""
        ref = getattr(self, '_loss_ref', None)
        if ref is None:
            raise RuntimeError('loss_value(): no training_step(reduce=True) has run yet')
        s = self._read_stream
        s.wait_event(self._loss_evt)
        with torch.cuda.stream(s):
            self._loss_pin.copy_(ref, non_blocking=True)
        s.synchronize()
        return float(self._loss_pin)

    def _dp_sum(self, *tensors) -> None:
        grp = getattr(self, '_dp_group', None)
        if grp is not None:
            import torch.distributed as dist
            for t in tensors:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=grp)

    # ---- hot path ---------------------------------------------------------------------------
    def _neg_sampling(self, batch):
        """One uniform negative per target slot, shape target.shape + (1,) (basemodel.py:50-61)."""
        tgt = batch[self.fiid]
        self._neg_step += 1
        rank = 0
        if getattr(self, '_dp_group', None) is not None:
            import torch.distributed as dist
            rank = dist.get_rank(self._dp_group)
        return _engine.neg_sample(tuple(tgt.shape) + (1,), self.num_items, self.config['train'].get('seed', 0) + 7919 * rank,
                                  self._neg_step, tgt.device)

    def _publish_grads(self) -> None:
        """Expose the kernel-written gradient buffers as .grad (what loss.backward() leaves behind).  The buffers are
        persistent, so after the first backward this is a two-pointer check."""
        if self._flat_params[-1].grad is self._grad_views[-1] and self._flat_params[0].grad is self._grad_views[0] \
                and self.item_embedding.weight.grad is self._table_grad:
            return
        for p, g in zip(self._flat_params, self._grad_views):
            p.grad = g
        self.item_embedding.weight.grad = self._table_grad

    def _table_grad_buffer(self) -> torch.Tensor:
        """Zeroed accumulator for the embedding gradient (the fused Adam clears it while reading)."""
        grp = self._table_group
        if grp is None or grp.dirty:
            self._table_grad.zero_()
        if grp is not None:
            grp.dirty = True
        return self._table_grad

    def _step_forward(self, batch, reduce, return_query):
        """-> (loss tensor, query or None, opaque state for _step_backward)"""
        raise NotImplementedError

    def _step_backward(self, state, reduce, dloss, dquery) -> None:
        raise NotImplementedError

    def training_step(self, batch, reduce=True, return_query=False):
        """Reference model/basemodel.py:204-214: loss (and query) of one batch; `.backward()` fills `.grad`."""
        return _TrainStep.apply(self.item_embedding.weight, self, batch, reduce, return_query)

    def _item_dead(self, domain: str) -> torch.Tensor:
        """u8 mask of ids outside the eval domain (always id 0), built once per domain and kept on
        the device (the reference re-uploads the python list every batch, basemodel.py:358-359)."""
        m = self._dead_cache.get(domain)
        if m is None or m.device != self.item_embedding.weight.device:
            m = torch.ones(self.num_items, dtype=torch.uint8)
            m[torch.as_tensor(list(self.domain_item_mapping[domain]), dtype=torch.int64)] = 0
            m = self._dead_cache[domain] = m.to(self.item_embedding.weight.device)
        return m

    @torch.no_grad()
    def topk(self, batch, k, user_h=None):
        query = self.forward(batch)
        dead = self._item_dead(self.eval_domain)
        sh = self._shard if self._shard is not None else getattr(self, '_peer_eval', None) if self._peer is not None else None
        if sh is not None:
            lo, hi = self._shard_rows
            return sh.topk(_engine.topk, query, self.item_embedding.weight.data, dead[lo:hi].contiguous(), user_h, k)
        return _engine.topk(query, self.item_embedding.weight.data, dead, user_h, k)

    def set_eval_domain(self, domain):
        self.eval_domain = domain

    # ---- epoch loops (host glue; mirrors basemodel.py:171-262) ------------------------------------
    def current_epoch_trainloaders(self, nepoch):
        return self.dataset_list[0].get_loader()

    def training_epoch(self, nepoch):
        outputs = []
        for batch in self.current_epoch_trainloaders(nepoch):
            batch = {k: v.to(self.device, non_blocking=True) for k, v in batch.items()}
            batch['neg_item'] = self._neg_sampling(batch)
            self.optimizer.zero_grad()
            loss = self.training_step(batch=batch)
            loss.backward()
            self.optimizer.step()
            outputs.append({'loss_0': loss.detach()})
        return [outputs]

    @torch.no_grad()
    def _eval_epoch(self, loader, cutoffs) -> Dict[str, float]:
        """ndcg@c / recall@c over the loader: the hit position and the metric numerators are one kernel per batch
        (dr4sr_rank_metrics) accumulating into a device buffer; the host reads it once per epoch."""
        cutoffs = [int(c) for c in cutoffs]
        sums = torch.zeros(2 * len(cutoffs), dtype=torch.float64, device=self.device)
        n = 0
        for batch in loader:
            batch = {k: v.to(self.device, non_blocking=True) for k, v in batch.items()}
            _, ids = self.topk(batch, self.config['eval']['topk'], batch['user_hist'])
            _engine.rank_metrics(ids, batch[self.fiid], cutoffs, sums)
            n += ids.size(0)
        host = (sums / max(n, 1)).tolist()
        out = {}
        for i, c in enumerate(cutoffs):
            out[f'ndcg@{c}'], out[f'recall@{c}'] = host[2 * i], host[2 * i + 1]
        return out

    def fit(self):
        self._init_model(self.dataset_list[0])
        self.fit_loop()

    def fit_loop(self):
        t, e = self.config['train'], self.config['eval']
        best, bad, self._best_state = -1.0, 0, None
        for epoch in range(t['epochs']):
            tic = time.time()
            self.train()
            outs = self.training_epoch(epoch)
            self.training_time += time.time() - tic
            tic = time.time()
            self.eval()
            self.logged_metrics = {'epoch': epoch}
            for domain in self.domain_name_list:
                val = self.dataset_list[1]
                val.set_eval_domain(domain)
                self.set_eval_domain(domain)
                m = self._eval_epoch(val.get_loader(), [e['cutoff'][0]])
                self.logged_metrics.update({f'{domain}_{k}': v for k, v in m.items()})
                for k, v in m.items():
                    self.logged_metrics[k] = self.logged_metrics.get(k, 0.0) + v
            self.inference_time += time.time() - tic
            self.logged_metrics['train_loss_0'] = float(torch.stack([o['loss_0'] for o in outs[0]]).mean())
            self.logger.info(self.logged_metrics)
            score = self.logged_metrics.get('ndcg@20', self.logged_metrics.get(f"ndcg@{e['cutoff'][0]}", 0.0))
            if score > best:
                best, bad = score, 0
                self._best_state = {k: v.detach().clone() for k, v in self.state_dict().items()}
                self._best_epoch = epoch
                self._best_metrics = dict(self.logged_metrics)      # the reference stores the best epoch's metrics (callbacks.py:70-76)
            else:
                bad += 1
                if bad >= t['early_stop_patience']:
                    break
        self.save_checkpoint()

    def save_checkpoint(self, path: Optional[str] = None) -> str:
        """Same dict layout as the reference's EarlyStopping.save_checkpoint (utils/callbacks.py:130-136)."""
        root = os.path.join(self.config['eval']['save_path'], type(self).__name__, self.config['data']['dataset'])
        os.makedirs(root, exist_ok=True)
        path = path or os.path.join(root, time.strftime('%Y-%m-%d-%H-%M-%S') + '.ckpt')
        state = self._best_state if getattr(self, '_best_state', None) is not None else self.state_dict()
        torch.save({'config': self.config, 'model': type(self).__name__, 'epoch': getattr(self, '_best_epoch', 0),
                    'parameters': state, 'metric': dict(getattr(self, '_best_metrics', None) or getattr(self, 'logged_metrics', {}))}, path)
        self.ckpt_path = path
        return path

    def evaluate(self) -> Dict:
        if self.ckpt_path:
            self.load_checkpoint(self.ckpt_path)
        self.eval()
        out: Dict[str, float] = {}
        test = self.dataset_list[-1]
        for domain in self.domain_name_list:
            test.set_eval_domain(domain)
            self.set_eval_domain(domain)
            m = self._eval_epoch(test.get_loader(), self.config['eval']['cutoff'])
            out.update({f'{domain}_{k}': v for k, v in m.items()})
            for k, v in m.items():
                out[k] = out.get(k, 0.0) + v
        self.logger.info(out)
        return out

    def load_checkpoint(self, path: str) -> None:
        ckpt = torch.load(path, map_location=self.device, weights_only=False)
        self.config = ckpt['config']
        self.load_state_dict(ckpt['parameters'])
