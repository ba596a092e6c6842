"""Shared helpers for the parity tests: golden-fixture loading and oracle construction."""
from __future__ import annotations

import os
from typing import Dict

import numpy as np
import torch

from oracle import dr4sr_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
ALIASES = ('query_encoder.item_encoder.weight', 'query_encoder.0.1.weight')


def load_fixture(name: str) -> Dict[str, Dict[str, torch.Tensor]]:
    """Returns {'param': {...}, 'batch': {...}, 'grad': {...}, ...} of torch tensors."""
    out: Dict[str, Dict[str, torch.Tensor]] = {}
    with np.load(os.path.join(GOLDEN, name)) as z:
        for key in z.files:
            group, _, leaf = key.partition('/')
            a = np.array(z[key])
            out.setdefault(group, {})[leaf] = str(a) if a.dtype.kind in 'US' else torch.from_numpy(a)
    return out


def load_params(model: torch.nn.Module, params: Dict[str, torch.Tensor]) -> torch.nn.Module:
    res = model.load_state_dict(params, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert set(res.missing_keys) <= set(ALIASES), res.missing_keys
    return model


def oracle_from_fixture(kind: str, fx) -> torch.nn.Module:
    p = fx['param']
    N, D = p['item_embedding.weight'].shape
    if kind == 'sasrec':
        F = p['query_encoder.transformer_layer.layers.0.linear1.weight'].shape[0]
        nl = 1 + max(int(k.split('.')[3]) for k in p if k.startswith('query_encoder.transformer_layer.layers.'))
        m = orc.OracleSASRec(N, embed_dim=D, hidden_size=F, layer_num=nl, dropout_rate=0.0)
    elif kind == 'gru4rec':
        H = p['query_encoder.0.3.gru.weight_hh_l0'].shape[1]
        m = orc.OracleGRU4Rec(N, embed_dim=D, hidden_size=H, dropout_rate=0.0)
    elif kind == 'fmlp':
        m = orc.OracleFMLP(N, embed_dim=D, dropout_rate=0.0)
    else:
        raise ValueError(kind)
    return load_params(m, p)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b| -- the scale-relative error used for fp32 parity."""
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# ---- measured parity errors: every fp comparison records (label, error, tolerance); tests/conftest.py prints the worst
# error per label in the terminal summary and writes gpurun_out/parity_errors.json, so the margins are known ----------
ERRLOG = []


def check_err(label: str, err: float, tol: float) -> None:
    ERRLOG.append((label, float(err), float(tol)))
    assert err < tol, f'{label}: error {err:.3e} >= tolerance {tol:.1e}'
