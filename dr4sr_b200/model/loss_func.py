"""Loss façade with the reference's class names (reference model/loss_func.py:5-49).

On the CUDA path the loss is not a separate module call: scoring, BCE and their backward are one
kernel (`dr4sr_score_bce`).  The classes exist so `config['model']['loss_fn']` resolves the same way
and so code that inspects `model.loss_fn` keeps working; they carry the kind only.
"""
from __future__ import annotations

import torch.nn as nn


class BinaryCrossEntropyLoss(nn.Module):
    kind = 'bce'

    def forward(self, pos_score, neg_score, reduce=True):
        raise RuntimeError('dr4sr_b200 computes the sampled BCE inside the fused scoring kernel; '
                           'call model.training_step(batch) (there is no PyTorch fallback).')


class BPRLoss(nn.Module):
    """`config['model']['loss_fn'] = 'bpr'`: -logsig(s+ - s-) averaged over the non-pad targets, one negative
    (model/loss_func.py:40-49).  The reference cannot reach it -- training_step passes `reduce` to this two-argument
    forward (TypeError at model/basemodel.py:210) -- so this is the intended loss, fixed: `dr4sr_score_loss` with
    DR4SR_LOSS_BPR computes it, and its gradients, in the same fused sweep as the BCE."""
    kind = 'bpr'

    def forward(self, pos_score, neg_score, reduce=True):
        raise RuntimeError('dr4sr_b200 computes the sampled BPR loss inside the fused scoring kernel; '
                           'call model.training_step(batch) (there is no PyTorch fallback).')
