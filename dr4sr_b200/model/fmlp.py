"""FMLP on libdr4sr (reference model/fmlp.py:8-51, module/layers.py:740-808).

The reference hard-codes max_seq_len 50, width 64 and dropout 0.5 (fmlp.py:11-13, layers.py:743-745,762);
here the width follows ``config['model']['embed_dim']`` (64 or 128) and the dropout follows
``config['model']['dropout_rate']`` (0.5 in configs/fmlp.yaml, i.e. the reference's effective value).
Parameter containers keep the reference's names, so ``state_dict()`` keys are identical.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import engine as _engine
from .basemodel import BaseModel


class FilterLayer(nn.Module):
    """Container for the learnable spectral filter (module/layers.py:740-759)."""

    def __init__(self, max_seq_len: int = 50, hidden_size: int = 64, dropout: float = 0.5) -> None:
        super().__init__()
        self.complex_weight = nn.Parameter(torch.randn(1, max_seq_len // 2 + 1, hidden_size, 2, dtype=torch.float32) * 0.02)
        self.out_dropout = nn.Dropout(dropout)
        self.LayerNorm = nn.LayerNorm(hidden_size, eps=1e-12)


class Intermediate(nn.Module):
    """Container for the position-wise FFN (module/layers.py:761-779)."""

    def __init__(self, hidden_size: int = 64, hidden_dropout_prob: float = 0.5) -> None:
        super().__init__()
        self.dense_1 = nn.Linear(hidden_size, hidden_size * 4)
        self.dense_2 = nn.Linear(4 * hidden_size, hidden_size)
        self.LayerNorm = nn.LayerNorm(hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(hidden_dropout_prob)


class Layer(nn.Module):
    def __init__(self, max_seq_len: int = 50, hidden_size: int = 64, dropout: float = 0.5) -> None:
        super().__init__()
        self.filterlayer = FilterLayer(max_seq_len, hidden_size, dropout)
        self.intermediate = Intermediate(hidden_size, dropout)


class FMLPEncoder(nn.Module):
    def __init__(self, num_hidden_layers: int = 2, max_seq_len: int = 50, hidden_size: int = 64, dropout: float = 0.5) -> None:
        super().__init__()
        first = Layer(max_seq_len, hidden_size, dropout)
        self.layer = nn.ModuleList([first] + [Layer(max_seq_len, hidden_size, dropout) for _ in range(num_hidden_layers - 1)])
        for blk in self.layer[1:]:          # the reference deep-copies ONE Layer (module/layers.py:796-798)
            blk.load_state_dict(first.state_dict())

    def forward(self, hidden_states, output_all_encoded_layers=True):
        raise RuntimeError('FMLPEncoder is a parameter container here; call FMLP.forward(batch)')


class FMLP(BaseModel):
    def __init__(self, config, dataset_list) -> None:
        super().__init__(config, dataset_list)
        p = config['model'].get('dropout_rate', 0.5)
        self.position_embeddings = nn.Embedding(self.max_seq_len, self.embed_dim)
        self.LayerNorm = nn.LayerNorm(self.embed_dim, eps=1e-12)
        self.dropout = nn.Dropout(p)
        self.item_encoder = FMLPEncoder(config['model']['layer_num'], self.max_seq_len, self.embed_dim, p)

    def _build_engine(self) -> None:
        m = self.config['model']
        if self._shard_rows is not None:
            # the FMLP kernels index the table with global ids; a row-sharded table would be read out of bounds
            raise _engine._lib.Dr4srError("FMLP does not support config['train']['table_shard'] (row-sharded item table); "
                                          'use the replicated data-parallel layout')
        self.engine = _engine.FMLPEngine(self.num_items, self.embed_dim, self.max_seq_len, m['layer_num'], m.get('dropout_rate', 0.5),
                                         self.config['train'].get('seed', 0), self.item_embedding.weight.device)

    def _flat_parameters(self):
        out = [self.position_embeddings.weight, self.LayerNorm.weight, self.LayerNorm.bias]
        for blk in self.item_encoder.layer:
            f, i = blk.filterlayer, blk.intermediate
            out += [f.complex_weight, f.LayerNorm.weight, f.LayerNorm.bias, i.dense_1.weight, i.dense_1.bias, i.dense_2.weight,
                    i.dense_2.bias, i.LayerNorm.weight, i.LayerNorm.bias]
        return out

    @torch.no_grad()
    def enable_sharded_table(self, group) -> None:
        raise _engine._lib.Dr4srError('FMLP does not support the row-sharded item table; use enable_data_parallel')

    def forward(self, batch, need_pooling=True):
        """Last position of the last layer, [B, D], in train and eval alike (model/fmlp.py:37-39)."""
        self._check_flat()
        eng = self.engine
        in_ids = batch['in_' + self.fiid]
        b = eng.prep(None, in_ids.size(0))
        eng.encode(b, self.item_embedding.weight.data, self._flat, in_ids, train=self.training)
        return b.q_last.clone()

    def _step_forward(self, batch, reduce, return_query):
        eng = self.engine
        table = self.item_embedding.weight.data
        in_ids, item_id = batch['in_' + self.fiid], batch[self.fiid]
        neg = batch['neg_item'].view(item_id.shape)
        b = eng.prep(item_id, in_ids.size(0))
        self._dp_sum(b.counts[1:2])
        if self.training:
            eng.step += 1
        eng.encode(b, table, self._flat, in_ids, train=self.training)
        eng.score_bce(b, table, item_id, neg, want_grad=False)
        loss = eng.reduce_loss(b) if reduce else b.loss_pos.clone()
        if reduce:
            self._dp_sum(loss)
            self._mark_loss(loss)
        return loss, (b.q_last.clone() if return_query else None), (b, in_ids, item_id, neg)

    def _step_backward(self, state, reduce, dloss, dquery) -> None:
        eng = self.engine
        b, in_ids, item_id, neg = state
        table = self.item_embedding.weight.data
        if reduce:
            eng.score_bce(b, table, item_id, neg, want_grad=True, upstream=dloss)
        else:
            eng.score_bce(b, table, item_id, neg, want_grad=True, loss_weight=dloss)
        if dquery is not None:
            b.dq += dquery
        eng.encode_bwd(b, table, self._flat, in_ids, self._flat_grad)
        tg = self._table_grad_buffer()
        eng.table_grad(b, in_ids, item_id, neg, tg, self._flat_grad[: eng.L * eng.D].view(eng.L, eng.D))
        if getattr(self, '_dp_group', None) is not None:
            self._reduce_grads(tg, None)

    def composite_forward(self, batch, table=None):
        """Twice-differentiable torch evaluation (MetaModel's outer step only), model/fmlp.py:18-39."""
        import torch.nn.functional as F
        ids = batch['in_' + self.fiid]
        L = ids.size(1)
        emb = self.item_embedding(ids) if table is None else torch.nn.functional.embedding(ids, table, padding_idx=0)
        x = emb + self.position_embeddings(torch.arange(L, device=ids.device)).unsqueeze(0)
        x = self.dropout(self.LayerNorm(x))
        for blk in self.item_encoder.layer:
            f, i = blk.filterlayer, blk.intermediate
            spec = torch.fft.rfft(x, dim=1, norm='ortho') * torch.view_as_complex(f.complex_weight)
            y = f.LayerNorm(f.out_dropout(torch.fft.irfft(spec, n=L, dim=1, norm='ortho')) + x)
            x = i.LayerNorm(i.dropout(i.dense_2(F.gelu(i.dense_1(y)))) + y)
        return x[:, -1]

    def training_step(self, batch, reduce=True, return_query=False, align=False):
        return super().training_step(batch, reduce, return_query)
