"""MetaModel (DR4SR+) on libdr4sr (reference model/metamodel.py:19-197, utils/utils.py:134-255).

Bi-level trainer around a sub-model.  What runs where:
  * inner step (every batch): the sub-model's kernels with ``reduce=False, return_query=True`` give the
    per-slot loss and the query; the 2-layer "personaliser" MLP + Gumbel-softmax that turns the query into
    per-slot weights is three tiny torch ops on the GPU (D -> D -> 2); the weighted loss' gradient flows
    back into the kernels as per-slot loss weights (``loss_weight`` of `dr4sr_score_bce`) and as a
    gradient on the query (added to dq before `dr4sr_*_bwd`).
  * outer step (every `interval` = 30 batches after warm-up): the implicit hypergradient needs the
    Hessian-vector products of the training loss through the encoder, i.e. double backward.  The kernels
    are first-order only, so this one step runs a differentiable composite of the SAME parameter tensors
    through the torch modules that hold them (`nn.TransformerEncoder` / `nn.GRU` containers, SDPA math
    backend) -- SURVEY.md section 7 step 9 / section 8f item 3 ("composite fallback first; native HVP next").
    It is 1 step in 30 and never touches the item-table optimizer path.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils.clip_grad import clip_grad_norm_

from ..utils.config import default_config
from .basemodel import BaseModel, normal_initialization


class Hypergrad:
    """Implicit differentiation with a truncated Neumann inverse-HVP (reference utils/utils.py:134-205)."""

    def __init__(self, learning_rate: float = 0.1, truncate_iter: int = 3) -> None:
        self.learning_rate, self.truncate_iter = learning_rate, truncate_iter

    def grad(self, loss_val, loss_train, aux_params, params):
        d_val = torch.autograd.grad(loss_val, params, retain_graph=True, allow_unused=True)
        d_train = torch.autograd.grad(loss_train, params, allow_unused=True, create_graph=True)
        keep = [i for i, (a, b) in enumerate(zip(d_val, d_train)) if a is not None and b is not None]
        d_val, d_train, params = [d_val[i] for i in keep], [d_train[i] for i in keep], [params[i] for i in keep]
        p = v = list(d_val)
        for _ in range(self.truncate_iter):
            hv = torch.autograd.grad(d_train, params, grad_outputs=v, retain_graph=True, allow_unused=True)
            hv = [torch.zeros_like(x) if h is None else h * self.learning_rate for h, x in zip(hv, v)]
            v = [a - b for a, b in zip(v, hv)]
            p = [a + b for a, b in zip(p, v)]
        v3 = torch.autograd.grad(d_train, aux_params, grad_outputs=p, allow_unused=True)
        return [None if g is None else -g for g in v3]


class MetaOptimizer:
    """Reference utils/utils.py:207-255."""

    def __init__(self, meta_optimizer, hpo_lr, truncate_iter=3, max_grad_norm=10) -> None:
        self.meta_optimizer = meta_optimizer
        self.hypergrad = Hypergrad(learning_rate=hpo_lr, truncate_iter=truncate_iter)
        self.max_grad_norm = max_grad_norm

    def step(self, train_loss, val_loss, parameters, aux_params, return_grads=False):
        self.meta_optimizer.zero_grad()
        grads = self.hypergrad.grad(loss_val=val_loss, loss_train=train_loss, aux_params=aux_params, params=parameters)
        for p, g in zip(aux_params, grads):
            p.grad = g
        if self.max_grad_norm is not None:
            clip_grad_norm_([p for p in aux_params if p.grad is not None], max_norm=self.max_grad_norm)
        self.meta_optimizer.step()
        return grads if return_grads else None


class MetaModel(BaseModel):
    def __init__(self, config: Dict, dataset_list: List) -> None:
        super().__init__(config, dataset_list)
        self.interval = config['train']['interval']
        self.step_counter = 0
        self.item_embedding = None                      # MetaModel is a trainer; the table belongs to the sub-model
        self.tau = nn.Parameter(torch.ones(1) * 10)
        self.counter = 0
        self._gumbel_override = None                    # tests inject the Gumbel noise here

    # ---- setup (metamodel.py:29-86) ---------------------------------------------------------------
    def _init_model(self, train_data=None):
        self.sub_model = self._register_sub_model()
        self.sub_model._init_model(train_data)
        self.item_embedding = self.sub_model.item_embedding
        self.engine = self.sub_model.engine
        dev = self.sub_model.item_embedding.weight.device
        self.tau.data = self.tau.data.to(dev)
        self.meta_module = self._register_meta_modules().to(dev)
        self.meta_module.apply(normal_initialization)
        self.meta_optimizer = self._get_meta_optimizers()
        self.optimizer = self.sub_model.optimizer

    def _register_sub_model(self) -> BaseModel:
        name = self.config['model']['sub_model']
        sub_cfg = default_config(name)
        sub_cfg['data'].update({k: v for k, v in self.config['data'].items()})
        sub_cfg['model']['embed_dim'] = self.config['model']['embed_dim']
        for k in ('dropout_rate',):
            if k in self.config['model'] and name != 'GRU4Rec':
                sub_cfg['model'][k] = self.config['model'][k]
        sub_cfg['train']['device'] = self.config['train']['device']
        sub_cfg['train']['seed'] = self.config['train'].get('seed', 0)
        sub_cfg['train']['batch_size'] = self.config['train']['batch_size']
        import importlib
        cls = getattr(importlib.import_module('dr4sr_b200.model.' + name.lower()), name)
        return cls(sub_cfg, self.dataset_list)

    def _register_meta_modules(self) -> nn.Module:
        return nn.Sequential(nn.Linear(self.embed_dim, self.embed_dim), nn.ReLU(), nn.Linear(self.embed_dim, 2))

    def _get_meta_optimizers(self):
        t = self.config['train']
        params = list(self.meta_module.parameters()) + [self.tau]
        name = t['meta_optimizer'].lower()
        if name == 'adam':
            opt = torch.optim.Adam(params, lr=t['meta_learning_rate'])
        elif name == 'sgd':
            opt = torch.optim.SGD(params, lr=t['meta_learning_rate'], weight_decay=t['meta_weight_decay'], momentum=0.9)
        else:
            opt = torch.optim.Adam(params, lr=t['meta_learning_rate'], weight_decay=t['meta_weight_decay'])
        return MetaOptimizer(opt, hpo_lr=t['hpo_learning_rate'])

    # ---- inner step (metamodel.py:169-194) --------------------------------------------------------
    def forward(self, batch):
        return self.sub_model.forward(batch)

    def selection(self, query):
        logits = self.meta_module(query)
        tau = torch.clip(self.tau, min=self.config['model']['tau_min'])
        if self._gumbel_override is not None:           # same formula as F.gumbel_softmax with the noise injected
            return F.softmax((logits + self._gumbel_override) / tau, dim=-1)[..., 0].squeeze()
        return F.gumbel_softmax(logits, tau=tau, dim=-1, hard=False)[..., 0].squeeze()

    def training_step(self, batch, reduce=True, return_query=True, align=False):
        loss_value, query = self.sub_model.training_step(batch, reduce=False, return_query=True)
        weight = self.selection(query)
        mask = batch['user_id'] == 0
        if weight.dim() == 2:
            mask = mask.unsqueeze(-1)
        weight = weight.masked_fill(mask, 1)
        weight = weight.masked_fill(batch[self.fiid] == 0, 0)
        self.counter += 1
        return (loss_value * weight).sum()

    # ---- epoch loop (metamodel.py:89-121) ----------------------------------------------------------
    def training_epoch(self, nepoch):
        outputs = []
        warm = self.config['train']['warmup_epoch']
        for batch in self.current_epoch_trainloaders(nepoch):
            batch = {k: v.to(self.device, non_blocking=True) for k, v in batch.items()}
            batch['neg_item'] = self.sub_model._neg_sampling(batch)
            self.sub_model.optimizer.zero_grad()
            loss = self.training_step(batch=batch) if nepoch > warm else self.sub_model.training_step(batch=batch)
            loss.backward()
            self.sub_model.optimizer.step()
            outputs.append({'loss_0': loss.detach()})
            self.step_counter += 1
            if self.step_counter % self.interval == 0 and nepoch > warm:
                self._outter_loop(nepoch)
        return [outputs]

    def current_epoch_metaloaders(self, nepoch):
        return self.dataset_list[0].get_loader()

    # ---- outer step (metamodel.py:123-166): composite, twice-differentiable evaluation ----------------
    def _composite_losses(self, batch, weighted: bool, table=None):
        """Per-slot BCE of the sub-model through the torch modules that hold its parameters (double-backward
        capable); same arithmetic as the reference's training_step (basemodel.py:204-214, loss_func.py:9-35)."""
        sm = self.sub_model
        q = sm.composite_forward(batch, table)
        E = sm.item_embedding.weight if table is None else table
        item_id, neg = batch[self.fiid], batch['neg_item']
        pos = (q * E[item_id]).sum(-1)
        negs = (q.unsqueeze(-2) * E[neg]).sum(-1)
        pad = item_id == 0
        n = (~pad).sum()
        per = (-F.logsigmoid(pos.masked_fill(pad, 0.0)).masked_fill(pad, 0.0) + F.softplus(negs).mean(-1).masked_fill(pad, 0.0)) / n
        if not weighted:
            return per.sum()
        weight = self.selection(q)
        mask = batch['user_id'] == 0
        if weight.dim() == 2:
            mask = mask.unsqueeze(-1)
        weight = weight.masked_fill(mask, 1).masked_fill(pad, 0)
        return (per * weight).sum()

    def _outer_step(self, val_batch, train_batch, return_grads=False):
        """One hypergradient update of the meta module from a validation batch and a training batch (both with
        `neg_item`): the else-branch of the reference's _outter_loop (metamodel.py:149-166)."""
        from torch.nn.attention import SDPBackend, sdpa_kernel
        # The item table enters both losses only through the rows the two batches touch: every other row has a zero gradient
        # and a zero row / column in the Hessian, so the hypergradient over the whole table (the reference differentiates all
        # N x D of it, several dense passes per Neumann term) equals the one over the touched rows.  Gather them once, remap
        # the ids (0 stays 0: the padding row), and differentiate with respect to that [rows, D] leaf instead.
        sm, fi = self.sub_model, self.fiid
        keys = ('in_' + fi, fi, 'neg_item')
        used = torch.cat([b[k].reshape(-1) for b in (val_batch, train_batch) for k in keys] + [torch.zeros(1, dtype=torch.int64, device=self.device)])
        rows = torch.unique(used)                        # sorted, rows[0] == 0
        sub_table = sm.item_embedding.weight.detach()[rows].requires_grad_(True)
        val_batch, train_batch = dict(val_batch), dict(train_batch)
        for b in (val_batch, train_batch):
            for k in keys:
                b[k] = torch.searchsorted(rows, b[k])
        table_ptr = sm.item_embedding.weight.data_ptr()
        others = [p for p in sm.parameters() if p.data_ptr() != table_ptr]
        with sdpa_kernel(SDPBackend.MATH), torch.backends.cudnn.flags(enabled=False):
            meta_loss = self._composite_losses(val_batch, weighted=False, table=sub_table)
            meta_train_loss = self._composite_losses(train_batch, weighted=True, table=sub_table)
            return self.meta_optimizer.step(val_loss=meta_loss, train_loss=meta_train_loss,
                                            aux_params=list(self.meta_module.parameters()),
                                            parameters=others + [sub_table], return_grads=return_grads)

    def _outter_loop(self, nepoch):
        def one_batch(loader):
            batch = next(iter(loader))
            batch = {k: v.to(self.device) for k, v in batch.items()}
            batch['neg_item'] = self.sub_model._neg_sampling(batch)
            return batch
        self._outer_step(one_batch(self.current_epoch_metaloaders(nepoch)), one_batch(self.current_epoch_trainloaders(nepoch)))

    # ---- delegate the rest to the sub-model ------------------------------------------------------------
    def topk(self, batch, k, user_h=None):
        self.sub_model.eval_domain = self.eval_domain
        return self.sub_model.topk(batch, k, user_h)

    def train(self, mode: bool = True):
        super().train(mode)
        if hasattr(self, 'sub_model'):
            self.sub_model.train(mode)
        return self
