"""dr4sr_oracle -- CPU (PyTorch fp32) restatement of DR4SR's sequential-recommender hot path.

TEST INFRASTRUCTURE.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module.  The product
(``dr4sr_b200``) never does: it fails loudly when its CUDA library is missing.

Parity status: PINNED.
  * ``tests/golden/make_golden.py`` imports the *unmodified* reference from /root/reference
    (this container only) and dumps inputs / parameters / outputs / gradients / post-Adam
    parameters; ``tests/test_oracle_golden.py`` checks this restatement against those dumps.
  * the shipped ``pre-trained_embedding.ckpt`` known answers (ndcg@20 / recall@20 stored in the
    checkpoint) are reproduced through this restatement.

The arithmetic of the reference lives in a third-party dependency (``torch``; requirements.txt:1,
README.md:18 says 1.13.1, this image has 2.11.0).  The module-style classes below therefore call
the *same* torch building blocks at the same call sites as the reference (faithful port, also the
honest CPU speed of the reference), while the ``*_explicit`` functions restate the published
algorithm of those building blocks in elementary tensor algebra so every kernel has an
independent spec.  Both are tested equal.

Every function cites the reference file:line (relative to /root/reference) it follows.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

Tensor = torch.Tensor
Batch = Dict[str, Tensor]


# ----------------------------------------------------------------------------------------------
# initialisation (utils/utils.py:70-81, model/basemodel.py:44-45)
# ----------------------------------------------------------------------------------------------
def reference_init_(model: nn.Module, std: float = 0.02) -> nn.Module:
    """``self.apply(normal_initialization)``: Embedding / Linear weights ~ N(0, std), pad row and
    Linear biases zero, LayerNorm (1, 0).  Everything else (MHA in_proj, GRU, FMLP complex
    weight) keeps its constructor default."""
    for m in model.modules():
        if isinstance(m, nn.Embedding):
            m.weight.data.normal_(0.0, std)
            if m.padding_idx is not None:
                m.weight.data[m.padding_idx].zero_()
        elif isinstance(m, nn.Linear):
            m.weight.data.normal_(0.0, std)
            if m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.LayerNorm):
            m.bias.data.zero_()
            m.weight.data.fill_(1.0)
    return model


# ----------------------------------------------------------------------------------------------
# negatives (model/basemodel.py:50-61)
# ----------------------------------------------------------------------------------------------
def neg_sampling(target: Tensor, num_items: int, generator: Optional[torch.Generator] = None) -> Tensor:
    """One negative per target slot, i.i.d. uniform on {1..N-1}, with replacement, positives not
    excluded (the exclusion is commented out at basemodel.py:54).  Shape = target.shape + (1,).
    The reference draws them with ``multinomial`` over a B x N weight matrix; the distribution
    is identical, the random stream is not (SURVEY.md section 7 'Negative sampling')."""
    return torch.randint(1, num_items, tuple(target.shape) + (1,), generator=generator, dtype=torch.int64)


def neg_sampling_as_reference(batch: Batch, num_items: int, max_seq_len: int) -> Tensor:
    """The literal reference procedure (B x N weight matrix + multinomial), kept for the CPU
    baseline timing so the baseline pays what the reference pays (basemodel.py:50-61)."""
    seq = batch['in_item_id']
    w = torch.ones(seq.shape[0], num_items)
    w[:, 0] = 0
    n = max_seq_len if batch['item_id'].dim() == 2 else 1
    idx = torch.multinomial(w, n, replacement=True)
    return idx.reshape_as(batch['item_id']).unsqueeze(-1)


# ----------------------------------------------------------------------------------------------
# loss (model/loss_func.py:5-49)
# ----------------------------------------------------------------------------------------------
def bce_loss(pos_score: Tensor, neg_score: Tensor, reduce: bool = True) -> Tensor:
    """BinaryCrossEntropyLoss.forward (loss_func.py:9-35).  ``pos_score`` carries -inf at pad
    targets; n = number of finite positives; negatives are averaged over the last dim."""
    pad = torch.isinf(pos_score)
    n = (~pad).sum()
    pos = F.logsigmoid(pos_score).masked_fill(pad, 0.0)
    neg = (F.softplus(neg_score) / neg_score.size(-1)).sum(-1)
    if pos_score.dim() == neg_score.dim() - 1:
        neg = neg.masked_fill(pad, 0.0)
        if reduce:
            return -(pos.sum() / n) + neg.sum() / n
        return -(pos / n) + neg / n
    # loss_func.py:33-34 -- unreachable from training_step (neg always has one more dim)
    pos = pos.sum() / n if reduce else pos / n
    return -pos + neg.mean()


def bpr_loss(pos_score: Tensor, neg_score: Tensor) -> Tensor:
    """BPRLoss.forward (loss_func.py:44-49): -sum_valid logsigmoid(s+ - s-) / n.  Unreachable via
    training_step in the reference (no ``reduce`` kwarg -> TypeError); restated as the intended
    semantics for the fixed extension."""
    pad = torch.isinf(pos_score)
    l = F.logsigmoid(pos_score.unsqueeze(-1) - neg_score).masked_fill(pad.unsqueeze(-1), 0.0)
    w = 1.0 / neg_score.size(-1)
    return -(l * w).sum(-1).sum() / (~pad).sum()


def sampled_scores(query: Tensor, table: Tensor, item_id: Tensor, neg_item: Tensor) -> Tuple[Tensor, Tensor]:
    """BaseModel.training_step scoring (basemodel.py:205-208)."""
    pos = (query * table[item_id]).sum(-1)
    neg = (query.unsqueeze(-2) * table[neg_item]).sum(-1)
    pos[item_id == 0] = -torch.inf          # in place, as the reference (keeps autograd's accumulation order)
    return pos, neg


# ----------------------------------------------------------------------------------------------
# pooling (module/layers.py:41-50, 69-73)
# ----------------------------------------------------------------------------------------------
def pool_origin(x: Tensor, seqlen: Tensor) -> Tensor:
    L = x.size(1)
    dead = torch.arange(L, device=x.device).view(1, L, 1) >= seqlen.view(-1, 1, 1)
    return x.masked_fill(dead, 0.0)


def pool_last(x: Tensor, seqlen: Tensor) -> Tensor:
    idx = (seqlen - 1).view(-1, 1, 1).expand(-1, 1, x.size(-1))
    return x.gather(1, idx).squeeze(1)


# ----------------------------------------------------------------------------------------------
# common trainer surface (model/basemodel.py)
# ----------------------------------------------------------------------------------------------
class _OracleBase(nn.Module):
    """Holds the shared item table and the scoring / eval functions of BaseModel."""

    def __init__(self, num_items: int, embed_dim: int, max_seq_len: int = 50) -> None:
        super().__init__()
        self.num_items, self.embed_dim, self.max_seq_len = num_items, embed_dim, max_seq_len
        self.item_embedding = nn.Embedding(num_items, embed_dim, padding_idx=0)   # basemodel.py:42

    def init_reference_style(self) -> "_OracleBase":
        return reference_init_(self)

    def training_step(self, batch: Batch, reduce: bool = True, return_query: bool = False):
        """basemodel.py:204-214."""
        q = self.forward(batch)
        pos, neg = sampled_scores(q, self.item_embedding.weight, batch['item_id'], batch['neg_item'])
        loss = bce_loss(pos, neg, reduce=reduce)
        return (loss, q) if return_query else loss

    @torch.no_grad()
    def topk(self, batch: Batch, k: int, domain_items: Optional[Sequence[int]] = None) -> Tuple[Tensor, Tensor]:
        """basemodel.py:354-365: full-catalog scores, -inf outside the eval domain (always
        includes id 0) and at the user's history ids, then top-k."""
        q = self.forward(batch)
        score = q @ self.item_embedding.weight[: self.num_items].T
        dead = torch.ones(1, self.num_items, dtype=torch.bool, device=score.device)
        if domain_items is None:
            dead[:, 1:] = False
        else:
            dead[:, torch.as_tensor(list(domain_items), dtype=torch.int64)] = False
        score = score.masked_fill(dead, -torch.inf)
        score = torch.scatter(score, 1, batch['user_hist'], -torch.inf)
        return torch.topk(score, k)

    def make_adam(self, lr: float = 1e-3, weight_decay: float = 0.0) -> torch.optim.Optimizer:
        """basemodel.py:79-86 (dense Adam over every parameter, table included)."""
        return torch.optim.Adam(self.parameters(), lr=lr, weight_decay=weight_decay)


# ----------------------------------------------------------------------------------------------
# SASRec (model/sasrec.py:10-97)
# ----------------------------------------------------------------------------------------------
class _SASRecEncoder(nn.Module):
    def __init__(self, table: nn.Embedding, D: int, L: int, nhead: int, ffn: int, p: float, act: str,
                 eps: float, nlayer: int) -> None:
        super().__init__()
        self.item_encoder = table                                       # sasrec.py:16 (shared table)
        self.position_emb = nn.Embedding(L, D)                          # sasrec.py:20
        block = nn.TransformerEncoderLayer(d_model=D, nhead=nhead, dim_feedforward=ffn, dropout=p,
                                           activation=act, layer_norm_eps=eps, batch_first=True,
                                           norm_first=False)             # sasrec.py:21-30
        self.transformer_layer = nn.TransformerEncoder(block, num_layers=nlayer)   # sasrec.py:31-34
        self.dropout = nn.Dropout(p)


class OracleSASRec(_OracleBase):
    def __init__(self, num_items: int, embed_dim: int = 64, max_seq_len: int = 50, head_num: int = 2,
                 hidden_size: int = 128, dropout_rate: float = 0.5, activation: str = 'gelu',
                 layer_norm_eps: float = 1e-12, layer_num: int = 2) -> None:
        super().__init__(num_items, embed_dim, max_seq_len)
        self.query_encoder = _SASRecEncoder(self.item_embedding, embed_dim, max_seq_len, head_num, hidden_size,
                                            dropout_rate, activation, layer_norm_eps, layer_num)

    def encode(self, batch: Batch) -> Tensor:
        """sasrec.py:39-68 without pooling."""
        enc = self.query_encoder
        ids = batch['in_item_id']
        L = ids.size(1)
        x = enc.item_encoder(ids) + enc.position_emb(torch.arange(L, device=ids.device)).unsqueeze(0)      # :42-46,64
        causal = torch.triu(torch.ones(L, L, dtype=torch.bool, device=ids.device), 1)                       # :58
        return enc.transformer_layer(src=enc.dropout(x), mask=causal, src_key_padding_mask=ids == 0)  # :65-68

    def forward(self, batch: Batch) -> Tensor:
        out = self.encode(batch)
        return pool_origin(out, batch['seqlen']) if self.training else pool_last(out, batch['seqlen'])  # :72-75


def layer_norm_explicit(z: Tensor, g: Tensor, b: Tensor, eps: float) -> Tensor:
    mu = z.mean(-1, keepdim=True)
    var = ((z - mu) ** 2).mean(-1, keepdim=True)
    return (z - mu) * torch.rsqrt(var + eps) * g + b


def gelu_erf(x: Tensor) -> Tensor:
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def sasrec_layer_explicit(x: Tensor, key_is_pad: Tensor, w: Dict[str, Tensor], nhead: int, eps: float) -> Tensor:
    """One post-norm encoder layer, dropout 0 (torch's
    TransformerEncoderLayer as configured at sasrec.py:21-30).  ``w`` uses the state_dict leaf
    names of ``transformer_layer.layers.{i}``."""
    B, L, D = x.shape
    dh = D // nhead
    qkv = x @ w['self_attn.in_proj_weight'].T + w['self_attn.in_proj_bias']
    q, k, v = (t.view(B, L, nhead, dh).transpose(1, 2) for t in qkv.split(D, dim=-1))
    s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    dead = torch.triu(torch.ones(L, L, dtype=torch.bool, device=x.device), 1).view(1, 1, L, L) | key_is_pad.view(B, 1, 1, L)
    a = torch.softmax(s.masked_fill(dead, -torch.inf), dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, L, D)
    o = o @ w['self_attn.out_proj.weight'].T + w['self_attn.out_proj.bias']
    x1 = layer_norm_explicit(x + o, w['norm1.weight'], w['norm1.bias'], eps)
    h = gelu_erf(x1 @ w['linear1.weight'].T + w['linear1.bias'])
    f = h @ w['linear2.weight'].T + w['linear2.bias']
    return layer_norm_explicit(x1 + f, w['norm2.weight'], w['norm2.bias'], eps)


def sasrec_encode_explicit(model: OracleSASRec, batch: Batch) -> Tensor:
    """Whole encoder in elementary algebra, dropout 0 (kernel spec for K1 + K2)."""
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    ids = batch['in_item_id']
    L = ids.size(1)
    x = sd['item_embedding.weight'][ids] + sd['query_encoder.position_emb.weight'][:L].unsqueeze(0)
    nlayer = len(model.query_encoder.transformer_layer.layers)
    blk = model.query_encoder.transformer_layer.layers[0]
    for i in range(nlayer):
        pre = f'query_encoder.transformer_layer.layers.{i}.'
        w = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
        x = sasrec_layer_explicit(x, ids == 0, w, blk.self_attn.num_heads, blk.norm1.eps)
    return x


# ----------------------------------------------------------------------------------------------
# GRU4Rec (model/gru4rec.py:9-31, module/layers.py:117-136)
# ----------------------------------------------------------------------------------------------
class _Pick(nn.Module):
    def forward(self, batch: Batch) -> Tensor:           # LambdaLayer(lambda x: x['in_item_id'])
        return batch['in_item_id']


class _GRUWrap(nn.Module):
    def __init__(self, D: int, H: int, nlayer: int) -> None:
        super().__init__()
        self.gru = nn.GRU(input_size=D, hidden_size=H, num_layers=nlayer, bias=False, batch_first=True)

    def forward(self, x: Tensor) -> Tensor:
        return self.gru(x)[0]


class OracleGRU4Rec(_OracleBase):
    def __init__(self, num_items: int, embed_dim: int = 64, max_seq_len: int = 50, hidden_size: int = 256,
                 dropout_rate: float = 0.2, layer_num: int = 2) -> None:
        super().__init__(num_items, embed_dim, max_seq_len)
        # nesting reproduces the reference's state_dict keys query_encoder.0.{1,3}.*, query_encoder.1.*
        self.query_encoder = nn.Sequential(
            nn.Sequential(_Pick(), self.item_embedding, nn.Dropout(dropout_rate),
                          _GRUWrap(embed_dim, hidden_size, layer_num)),          # gru4rec.py:14-19
            nn.Linear(hidden_size, embed_dim))                                    # gru4rec.py:20

    def encode(self, batch: Batch) -> Tensor:
        return self.query_encoder(batch)

    def forward(self, batch: Batch) -> Tensor:
        out = self.encode(batch)
        return pool_origin(out, batch['seqlen']) if self.training else pool_last(out, batch['seqlen'])


def gru_explicit(x: Tensor, w_ih: Sequence[Tensor], w_hh: Sequence[Tensor]) -> Tensor:
    """Bias-free multi-layer GRU, h0 = 0, gate rows ordered [r; z; n]."""
    B, L, _ = x.shape
    for wi, wh in zip(w_ih, w_hh):
        H = wh.size(1)
        h = x.new_zeros(B, H)
        gi_all = x @ wi.T
        outs = []
        for t in range(L):
            gi, gh = gi_all[:, t], h @ wh.T
            r = torch.sigmoid(gi[:, :H] + gh[:, :H])
            z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
            n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
            h = (1.0 - z) * n + z * h
            outs.append(h)
        x = torch.stack(outs, 1)
    return x


# ----------------------------------------------------------------------------------------------
# FMLP (model/fmlp.py:8-39, module/layers.py:740-808), width parametrised
# ----------------------------------------------------------------------------------------------
class _Filter(nn.Module):
    def __init__(self, L: int, D: int, p: float) -> None:
        super().__init__()
        self.complex_weight = nn.Parameter(torch.randn(1, L // 2 + 1, D, 2) * 0.02)   # layers.py:743
        self.out_dropout = nn.Dropout(p)
        self.LayerNorm = nn.LayerNorm(D, eps=1e-12)

    def forward(self, x: Tensor) -> Tensor:                                             # layers.py:747-759
        L = x.size(1)
        X = torch.fft.rfft(x, dim=1, norm='ortho') * torch.view_as_complex(self.complex_weight)
        f = torch.fft.irfft(X, n=L, dim=1, norm='ortho')
        return self.LayerNorm(self.out_dropout(f) + x)


class _Inter(nn.Module):
    def __init__(self, D: int, p: float) -> None:
        super().__init__()
        self.dense_1 = nn.Linear(D, 4 * D)                                              # layers.py:764
        self.dense_2 = nn.Linear(4 * D, D)
        self.LayerNorm = nn.LayerNorm(D, eps=1e-12)
        self.dropout = nn.Dropout(p)

    def forward(self, x: Tensor) -> Tensor:                                             # layers.py:771-779
        h = self.dense_2(F.gelu(self.dense_1(x)))
        return self.LayerNorm(self.dropout(h) + x)


class _FMLPBlock(nn.Module):
    def __init__(self, L: int, D: int, p: float) -> None:
        super().__init__()
        self.filterlayer = _Filter(L, D, p)
        self.intermediate = _Inter(D, p)

    def forward(self, x: Tensor) -> Tensor:
        return self.intermediate(self.filterlayer(x))


class _FMLPStack(nn.Module):
    def __init__(self, L: int, D: int, p: float, nlayer: int) -> None:
        super().__init__()
        self.layer = nn.ModuleList([_FMLPBlock(L, D, p) for _ in range(nlayer)])


class OracleFMLP(_OracleBase):
    """The reference hard-codes L=50, D=64, p=0.5 (fmlp.py:11-13, layers.py:743-745,762); this
    restatement takes them as arguments (D=128 is BASELINE config 3)."""

    def __init__(self, num_items: int, embed_dim: int = 64, max_seq_len: int = 50, layer_num: int = 2,
                 dropout_rate: float = 0.5) -> None:
        super().__init__(num_items, embed_dim, max_seq_len)
        self.position_embeddings = nn.Embedding(max_seq_len, embed_dim)    # fmlp.py:11
        self.LayerNorm = nn.LayerNorm(embed_dim, eps=1e-12)                # fmlp.py:12
        self.dropout = nn.Dropout(dropout_rate)                            # fmlp.py:13
        self.item_encoder = _FMLPStack(max_seq_len, embed_dim, dropout_rate, layer_num)
        # reference: all layers are deep copies of one Layer() => identical initial complex_weight
        for blk in self.item_encoder.layer[1:]:
            blk.load_state_dict(self.item_encoder.layer[0].state_dict())

    def encode(self, batch: Batch) -> Tensor:
        ids = batch['in_item_id']
        L = ids.size(1)
        x = self.item_embedding(ids) + self.position_embeddings(torch.arange(L, device=ids.device)).unsqueeze(0)   # fmlp.py:18-24
        x = self.dropout(self.LayerNorm(x))                                                    # fmlp.py:25-26
        for blk in self.item_encoder.layer:
            x = blk(x)
        return x

    def forward(self, batch: Batch) -> Tensor:
        return self.encode(batch)[:, -1]                                                       # fmlp.py:37-39


def fmlp_filter_explicit(x: Tensor, complex_weight: Tensor) -> Tensor:
    """rfft(ortho) * W -> irfft(ortho) as a per-channel circular convolution (Appendix C.4):
    f[t,d] = sum_s h_d[(t-s) mod L] x[s,d], h_d = irfft(W[:,d], n=L, norm='backward')."""
    B, L, D = x.shape
    W = torch.view_as_complex(complex_weight.contiguous())[0]              # [L//2+1, D]
    h = torch.fft.irfft(W, n=L, dim=0, norm='backward')                    # [L, D]
    idx = ((torch.arange(L).view(L, 1) - torch.arange(L).view(1, L)) % L).to(x.device)    # [t, s]
    K = h[idx]                                                             # [t, s, D]
    return torch.einsum('tsd,bsd->btd', K, x)


# ----------------------------------------------------------------------------------------------
# Adam (torch.optim.Adam as called at basemodel.py:85-86)
# ----------------------------------------------------------------------------------------------
def adam_explicit(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float = 1e-3, b1: float = 0.9,
                  b2: float = 0.999, eps: float = 1e-8, wd: float = 0.0) -> Tuple[Tensor, Tensor, Tensor]:
    """Single-tensor Adam exactly in torch's op order (torch/optim/adam.py _single_tensor_adam):
    g += wd*p; m = lerp(m, g, 1-b1); v = b2*v + (1-b2) g^2;
    p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)."""
    if wd != 0.0:
        g = g + wd * p
    m = m + (g - m) * (1.0 - b1)
    v = v * b2 + (1.0 - b2) * g * g
    bc1 = 1.0 - b1 ** step
    bc2_sqrt = math.sqrt(1.0 - b2 ** step)
    denom = v.sqrt() / bc2_sqrt + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


# ----------------------------------------------------------------------------------------------
# eval metrics for the known-answer tests (evaluation/__init__.py:9-36,107-134; single target)
# ----------------------------------------------------------------------------------------------
def hit_matrix(topk_ids: Tensor, target: Tensor) -> Tensor:
    return target.view(-1, 1) == topk_ids                                  # basemodel.py:343


def recall_at_k(hit: Tensor, k: int) -> Tensor:
    return hit[:, :k].sum(-1).float()                                      # one relevant item per user


def ndcg_at_k(hit: Tensor, k: int) -> Tensor:
    denom = torch.log2(torch.arange(k, dtype=torch.float32) + 2.0).view(1, -1)
    return (hit[:, :k].float() / denom).sum(-1)                            # ideal DCG = 1


# ----------------------------------------------------------------------------------------------
# MetaModel weighting (model/metamodel.py:52-57,169-194), noise injected for determinism
# ----------------------------------------------------------------------------------------------
def meta_weighted_loss(per_pos_loss: Tensor, query: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor,
                       tau: Tensor, tau_min: float, gumbel: Tensor, user_id: Tensor, item_id: Tensor) -> Tensor:
    """loss = sum(l * w), w = softmax((MLP(q) + G)/max(tau, tau_min))[..., 0], forced to 1 on
    pattern rows (user_id == 0) and 0 on pad targets.  ``gumbel`` is the -log(Exp(1)) noise
    F.gumbel_softmax would draw (metamodel.py:169-172)."""
    logits = F.relu(query @ w1.T + b1) @ w2.T + b2
    w = torch.softmax((logits + gumbel) / torch.clip(tau, min=tau_min), dim=-1)[..., 0]
    mask = user_id == 0
    if w.dim() == 2:
        mask = mask.unsqueeze(-1)
    w = w.masked_fill(mask, 1.0).masked_fill(item_id == 0, 0.0)
    return (per_pos_loss * w).sum()
