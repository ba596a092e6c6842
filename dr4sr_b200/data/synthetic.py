"""Seeded synthetic batches with the reference's batch contract.

Keys and dtypes follow ``SeparateDataset.__getitem__`` (reference data/dataset.py:149-164):
all int64; ``in_item_id``/``item_id`` post-padded with 0 and target-shifted for SASRec/GRU4Rec,
pre-padded with a single target id per row for FMLP (reference README.md:78).
"""
from __future__ import annotations

from typing import Dict

import torch


def synthetic_batch(batch_size: int, max_seq_len: int, num_items: int, seed: int = 0, layout: str = 'post',
                    min_len: int = 1, with_neg: bool = True, eval_mode: bool = False, mean_len: float = 0.0) -> Dict[str, torch.Tensor]:
    """seqlen ~ U{min_len..L} (or, with mean_len > 0, 1 + Geometric with that mean, clipped to L: the shape of the shipped
    amazon-toys train_regen.pth, mean 2.25); ids ~ U{1..N-1}; negatives ~ U{1..N-1} (one per target slot)."""
    g = torch.Generator().manual_seed(seed)
    B, L, N = batch_size, max_seq_len, num_items
    if mean_len > 0.0:
        p = 1.0 / max(mean_len, 1.0 + 1e-6)                     # P(len = k) = p (1-p)^(k-1), k >= 1
        u = torch.rand(B, generator=g, dtype=torch.float64).clamp_min(1e-12)
        seqlen = (1 + torch.floor(torch.log(u) / torch.log(torch.tensor(1.0 - p, dtype=torch.float64)))).clamp(1, L).to(torch.int64)
    else:
        seqlen = torch.randint(min_len, L + 1, (B,), generator=g, dtype=torch.int64)
    x = torch.randint(1, N, (B, L + 1), generator=g, dtype=torch.int64)
    t = torch.arange(L).view(1, L)
    if layout == 'post':
        valid = t < seqlen.view(B, 1)
        in_ids = torch.where(valid, x[:, :L], torch.zeros_like(x[:, :L]))
        if eval_mode:
            target = x[:, L].clone()
        else:
            target = torch.where(valid, x[:, 1:], torch.zeros_like(x[:, 1:]))
    elif layout == 'pre':
        shift = (L - seqlen).view(B, 1)
        src = (t - shift).clamp(min=0)
        in_ids = torch.where(t >= shift, x[:, :L].gather(1, src), torch.zeros_like(src))
        target = x[:, L].clone()
    else:
        raise ValueError(f"layout must be 'post' or 'pre', got {layout!r}")
    batch = {
        'user_id': torch.randint(1, 1 << 20, (B,), generator=g, dtype=torch.int64),
        'in_item_id': in_ids,
        'item_id': target,
        'seqlen': seqlen,
        'label': torch.ones_like(target),
        'domain_id': torch.zeros_like(in_ids),
        'index': torch.arange(B, dtype=torch.int64),
    }
    if eval_mode:
        batch['user_hist'] = in_ids.clone()
    if with_neg:
        batch['neg_item'] = torch.randint(1, N, tuple(target.shape) + (1,), generator=g, dtype=torch.int64)
    return batch
