"""SASRec on libdr4sr (reference model/sasrec.py:10-120).

`SASRecQueryEncoder` keeps the reference's parameter containers (``item_encoder``, ``position_emb``,
``transformer_layer`` = torch.nn.TransformerEncoder) so ``state_dict()`` round-trips with the shipped
checkpoints, but none of their ``forward``s run: the arithmetic is `dr4sr_sasrec_fwd/bwd`,
`dr4sr_score_bce` and `dr4sr_table_grad` (include/dr4sr.h).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import engine as _engine
from .basemodel import BaseModel


class SASRecQueryEncoder(nn.Module):
    """Parameter container with the reference's attribute names (model/sasrec.py:11-37)."""

    def __init__(self, fiid, embed_dim, max_seq_len, n_head, hidden_size, dropout, activation, layer_norm_eps, n_layer,
                 item_encoder, bidirectional=False, training_pooling_type='origin', eval_pooling_type='last') -> None:
        super().__init__()
        if activation != 'gelu' or bidirectional or training_pooling_type != 'origin' or eval_pooling_type != 'last':
            raise NotImplementedError('libdr4sr implements the configuration the reference ships: causal, GELU, '
                                      "'origin'/'last' pooling (configs/sasrec.yaml)")
        self.fiid = fiid
        self.item_encoder = item_encoder
        self.position_emb = nn.Embedding(max_seq_len, embed_dim)
        block = nn.TransformerEncoderLayer(d_model=embed_dim, nhead=n_head, dim_feedforward=hidden_size, dropout=dropout,
                                           activation=activation, layer_norm_eps=layer_norm_eps, batch_first=True,
                                           norm_first=False)
        self.transformer_layer = nn.TransformerEncoder(encoder_layer=block, num_layers=n_layer)
        self.dropout = nn.Dropout(p=dropout)

    def flat_parameters(self):
        """Order of the C ABI's flat layout (include/dr4sr.h): positions, then per layer in state_dict order."""
        out = [self.position_emb.weight]
        for blk in self.transformer_layer.layers:
            out += [blk.self_attn.in_proj_weight, blk.self_attn.in_proj_bias, blk.self_attn.out_proj.weight,
                    blk.self_attn.out_proj.bias, blk.linear1.weight, blk.linear1.bias, blk.linear2.weight, blk.linear2.bias,
                    blk.norm1.weight, blk.norm1.bias, blk.norm2.weight, blk.norm2.bias]
        return out

    def forward(self, batch, need_pooling=True):
        raise RuntimeError('SASRecQueryEncoder is a parameter container here; call SASRec.forward(batch)')


def batch_len(b) -> torch.Tensor:
    return (b.tok_off[1:] - b.tok_off[:-1]).view(-1, 1)


class SASRec(BaseModel):
    def __init__(self, config, dataset_list) -> None:
        super().__init__(config, dataset_list)
        m = config['model']
        self.query_encoder = SASRecQueryEncoder(self.fiid, self.embed_dim, self.max_seq_len, m['head_num'], m['hidden_size'],
                                                m['dropout_rate'], m['activation'], m['layer_norm_eps'], m['layer_num'],
                                                self.item_embedding)

    def _build_engine(self) -> None:
        m = self.config['model']
        self.engine = _engine.SASRecEngine(self.num_items, self.embed_dim, self.max_seq_len, m['hidden_size'], m['head_num'],
                                           m['layer_num'], m['dropout_rate'], m['layer_norm_eps'],
                                           self.config['train'].get('seed', 0), self.item_embedding.weight.device)

    def _flat_parameters(self):
        return self.query_encoder.flat_parameters()

    @torch.no_grad()
    def forward(self, batch, need_pooling=True):
        """Eval: row seqlen-1 ('last' pooling) [B, D]; train: zero-padded rows ('origin') [B, L, D]."""
        self._check_flat()
        eng = self.engine
        in_ids = batch['in_' + self.fiid]
        b = eng.prep(batch['seqlen'], None)
        if self._peer is not None:
            self._peer.barrier()              # rows are read from the other ranks' shards: every rank's last update is done
        table, in_ids, _, _ = self._rows_for(b, in_ids, None, None)
        if self.training or not need_pooling:
            q_dense = torch.empty(in_ids.size(0), eng.L, eng.D, dtype=torch.float32, device=in_ids.device)
            eng.encode(b, table, self._flat, in_ids, train=self.training, q_dense=q_dense)
            return q_dense
        eng.encode(b, table, self._flat, in_ids, train=False, want_last=True)
        return b.q_last.clone()

    def _step_forward(self, batch, reduce, return_query):
        eng = self.engine
        in_ids, item_id, neg = batch['in_' + self.fiid], batch[self.fiid], batch['neg_item']
        neg = neg.view(item_id.shape)
        # peer-sharded table: rows are read from the other ranks' shards, so every rank's Adam pass of the previous step must be
        # done first.  The barrier is issued before the rank-local batch preparation and runs beside it on NCCL's stream.
        coll = self._peer is not None and self._peer.coll
        tick = self._peer.barrier_async() if (self._peer is not None and not coll) else None
        b = eng.prep(batch['seqlen'], item_id)
        if coll:                                       # one kernel over peer memory: the barrier + the global number of valid targets
            self._peer.barrier(count=b.counts[1:2])
            n_work = None
        else:
            n_work = self._dp_count_async(b.counts[1:2])   # data parallel: normalise by the global number of valid targets (needed
            if tick is not None:                           # by the loss kernel only: its all-reduce runs under the encoder forward)
                tick.wait()
        if self.training:
            eng.step += 1
        q_dense = None
        if return_query:
            q_dense = torch.empty(in_ids.size(0), eng.L, eng.D, dtype=torch.float32, device=in_ids.device)
        table, in_ids, item_id, neg = self._rows_for(b, in_ids, item_id, neg)     # sharded table: staged rows + remapped ids
        eng.encode(b, table, self._flat, in_ids, train=self.training, q_dense=q_dense)
        # training with the mean loss: ds, dq come out of the same sweep as the loss (scaled by autograd's grad_output
        # in the backward); the weighted / per-position variants recompute them there
        fused_grad = bool(reduce and self.training)
        if n_work is not None:
            n_work.wait()                     # (stream-side wait: the count's all-reduce ran under the encoder forward)
        eng.score_bce(b, table, item_id, neg, want_grad=fused_grad)
        loss = eng.reduce_loss(b) if reduce else b.loss_pos.clone()
        # data parallel, mean loss in training: the rank-local partial becomes the global loss inside the backward's single
        # gradient all-reduce (BaseModel._reduce_grads); any other use sums it right away
        late = bool(reduce and fused_grad and getattr(self, '_dp_group', None) is not None)
        if reduce and not late:
            self._dp_sum(loss)
        if reduce:
            self._mark_loss(loss, final=not late)
        # (a detached alias of the same storage: the returned tensor itself would close the cycle loss -> grad_fn -> ctx.state ->
        # loss, and the step's tensors would then wait for the cyclic GC instead of being freed by reference count -- measured:
        # sporadic cudaMallocs of the growing pool, 2 - 150 ms each with peer mappings enabled)
        return loss, q_dense, (b, table, in_ids, item_id, neg, fused_grad, loss.detach() if late else None)

    def _step_backward(self, state, reduce, dloss, dquery) -> None:
        eng = self.engine
        b, table, in_ids, item_id, neg, fused_grad, late_loss = state
        if fused_grad:
            eng.scale_grads(b, dloss)
        elif reduce:
            eng.score_bce(b, table, item_id, neg, want_grad=True, upstream=dloss)
        else:
            eng.score_bce(b, table, item_id, neg, want_grad=True, loss_weight=dloss)
        if dquery is not None:
            valid = torch.arange(eng.L, device=dquery.device).view(1, -1) < batch_len(b)
            b.dq[: int(b.counts[0])] += dquery[valid]
        # the target / negative rows of the embedding gradient need ds and q only: their scatter-add (NVLink atomics into the
        # owners' shards when the table is peer-sharded) runs on the background stream under the whole encoder backward; the
        # weight gradients finish on the library's side stream while the input rows are scattered here
        tg = self._scatter_target()
        split = not eng.deterministic_scatter          # (the sorted, bit-reproducible reduction takes all three row sets at once)
        if split:
            eng.table_grad_targets_async(b, item_id, neg, tg)
        gout = self._grad_out()
        eng.encode_bwd(b, table, self._flat, in_ids, gout, defer_join=True)
        pos = gout[: eng.L * eng.D].view(eng.L, eng.D)
        if split:
            eng.table_grad(b, in_ids, None, None, tg, pos)
            eng.table_grad_join()
        else:
            eng.table_grad(b, in_ids, item_id, neg, tg, pos)
        eng.join_bwd()
        if getattr(self, '_dp_group', None) is not None:
            self._reduce_grads(tg, late_loss)

    def composite_forward(self, batch, table=None):
        """Twice-differentiable torch evaluation of the same parameters (MetaModel's outer step only):
        the reference's own module graph, model/sasrec.py:39-75, 'origin' pooling.  `table`: rows to look the ids up in
        instead of the item table (MetaModel passes the touched rows with remapped ids)."""
        enc = self.query_encoder
        ids = batch['in_' + self.fiid]
        L = ids.size(1)
        ar = torch.arange(L, device=ids.device)
        emb = enc.item_encoder(ids) if table is None else torch.nn.functional.embedding(ids, table, padding_idx=0)
        x = emb + enc.position_emb(ar).unsqueeze(0)
        causal = torch.triu(torch.ones(L, L, dtype=torch.bool, device=ids.device), 1)
        out = enc.transformer_layer(src=enc.dropout(x), mask=causal, src_key_padding_mask=ids == 0)
        return out.masked_fill(ar.view(1, L, 1) >= batch['seqlen'].view(-1, 1, 1), 0.0)

    def training_step(self, batch, reduce=True, return_query=False, align=False):
        if align:
            raise NotImplementedError('the align branch (model/sasrec.py:111-119) has no caller in the reference')
        return super().training_step(batch, reduce, return_query)
