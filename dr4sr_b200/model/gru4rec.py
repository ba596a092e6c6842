"""GRU4Rec on libdr4sr (reference model/gru4rec.py:9-36)."""
from __future__ import annotations

import torch

from .. import engine as _engine
from ..module.layers import GRULayer, LambdaLayer, SeqPoolingLayer, VStackLayer
from .basemodel import BaseModel
from .sasrec import SASRec


class GRU4Rec(BaseModel):
    def __init__(self, config, dataset_list) -> None:
        super().__init__(config, dataset_list)
        m = self.config['model']
        # same nesting as the reference => same state_dict keys (query_encoder.0.{1,3}.*, query_encoder.1.*)
        self.query_encoder = VStackLayer(
            torch.nn.Sequential(
                LambdaLayer(lambda x: x['in_' + self.fiid]),
                self.item_embedding,
                torch.nn.Dropout(m['dropout_rate']),
                GRULayer(self.embed_dim, m['hidden_size'], m['layer_num']),
            ),
            torch.nn.Linear(m['hidden_size'], self.embed_dim))
        self.training_pooling_layer = SeqPoolingLayer(pooling_type='origin')
        self.eval_pooling_layer = SeqPoolingLayer(pooling_type='last')

    def _build_engine(self) -> None:
        m = self.config['model']
        self.engine = _engine.GRUEngine(self.num_items, self.embed_dim, self.max_seq_len, m['hidden_size'], m['layer_num'],
                                        m['dropout_rate'], self.config['train'].get('seed', 0), self.item_embedding.weight.device)

    def _flat_parameters(self):
        gru = self.query_encoder[0][3].gru
        out = []
        for l in range(gru.num_layers):
            out += [getattr(gru, f'weight_ih_l{l}'), getattr(gru, f'weight_hh_l{l}')]
        lin = self.query_encoder[1]
        return out + [lin.weight, lin.bias]

    # forward / training-step kernels are the packed-row flow of SASRec with the GRU engine underneath
    forward = SASRec.forward
    _step_forward = SASRec._step_forward

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
            from .sasrec import batch_len
            valid = torch.arange(eng.L, device=dquery.device).view(1, -1) < batch_len(b)
            b.dq[: int(b.counts[0])] += dquery[valid]
        eng.encode_bwd(b, table, self._flat, in_ids, self._flat_grad)
        tg = self._scatter_target()
        eng.table_grad(b, in_ids, item_id, neg, tg, None)          # no positional table in GRU4Rec
        if getattr(self, '_dp_group', None) is not None:
            self._reduce_grads(tg, late_loss)

    def composite_forward(self, batch, table=None):
        """Twice-differentiable torch evaluation (MetaModel's outer step only), model/gru4rec.py:23-31."""
        ids = batch['in_' + self.fiid]
        seq = self.query_encoder[0]
        emb = self.item_embedding(ids) if table is None else torch.nn.functional.embedding(ids, table, padding_idx=0)
        h = seq[3].gru(seq[2](emb))[0]
        out = self.query_encoder[1](h)
        ar = torch.arange(ids.size(1), device=ids.device)
        return out.masked_fill(ar.view(1, -1, 1) >= batch['seqlen'].view(-1, 1, 1), 0.0)

    def training_step(self, batch, reduce=True, return_query=False, align=False):
        return super().training_step(batch, reduce, return_query)
