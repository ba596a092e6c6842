"""CPU tests: the oracle restatement vs the golden vectors dumped from the unmodified reference
(tests/golden/make_golden.py) and vs its own elementary-algebra spec functions."""
import math

import pytest
import torch

from oracle import dr4sr_oracle as orc
from tests.helpers import load_fixture, oracle_from_fixture, rel_err, load_params

CASES = [('sasrec', 'sasrec_d64.npz'), ('sasrec', 'sasrec_d128.npz'), ('gru4rec', 'gru4rec_d64.npz'),
         ('gru4rec', 'gru4rec_d128.npz'), ('fmlp', 'fmlp_d64.npz')]


@pytest.mark.parametrize('kind,name', CASES)
def test_forward_loss_grads_match_reference(kind, name):
    fx = load_fixture(name)
    m = oracle_from_fixture(kind, fx).train()
    loss, q = m.training_step(fx['batch'], reduce=True, return_query=True)
    assert torch.allclose(q, fx['train']['query'], rtol=0, atol=1e-6)
    assert abs(float(loss) - float(fx['train']['loss'])) < 1e-6
    per = m.training_step(fx['batch'], reduce=False)
    assert torch.allclose(per, fx['train']['loss_per_pos'], rtol=1e-5, atol=1e-9)
    loss.backward()
    for k, p in m.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        assert rel_err(g, fx['grad'][k]) < 1e-5, k


@pytest.mark.parametrize('kind,name', CASES)
def test_adam_steps_match_reference(kind, name):
    fx = load_fixture(name)
    m = oracle_from_fixture(kind, fx).train()
    opt = m.make_adam(lr=float(fx['adam']['lr']), weight_decay=float(fx['adam']['weight_decay']))
    for want in fx['adam']['losses'].tolist():
        opt.zero_grad()
        loss = m.training_step(fx['batch'])
        loss.backward()
        opt.step()
        assert abs(float(loss) - want) < 2e-6
    for k, p in m.named_parameters():
        assert torch.allclose(p, fx['param_after'][k], rtol=0, atol=5e-6), k


@pytest.mark.parametrize('kind,name', CASES)
def test_eval_topk_matches_reference(kind, name):
    fx = load_fixture(name)
    m = oracle_from_fixture(kind, fx)
    load_params(m, fx['param_after']).eval()
    ev = fx['evalbatch']
    k = fx['eval']['topk_ids'].shape[1]
    s, i = m.topk(ev, k, None)
    assert torch.equal(i, fx['eval']['topk_ids'])
    assert torch.allclose(s, fx['eval']['topk_scores'], rtol=0, atol=1e-6)


def test_sasrec_explicit_spec_equals_module():
    fx = load_fixture('sasrec_d64.npz')
    m = oracle_from_fixture('sasrec', fx).train()
    with torch.no_grad():
        a = m.encode(fx['batch'])
        b = orc.sasrec_encode_explicit(m, fx['batch'])
    valid = torch.arange(50).view(1, -1) < fx['batch']['seqlen'].view(-1, 1)
    assert float((a - b)[valid].abs().max()) < 2e-6


def test_gru_explicit_spec_equals_module():
    fx = load_fixture('gru4rec_d64.npz')
    m = oracle_from_fixture('gru4rec', fx).train()
    gru = m.query_encoder[0][3].gru
    x = m.item_embedding(fx['batch']['in_item_id']).detach()
    with torch.no_grad():
        a = gru(x)[0]
        b = orc.gru_explicit(x, [gru.weight_ih_l0, gru.weight_ih_l1], [gru.weight_hh_l0, gru.weight_hh_l1])
    assert float((a - b).abs().max()) < 1e-6


def test_fmlp_filter_explicit_spec_equals_fft():
    torch.manual_seed(0)
    f = orc._Filter(50, 64, 0.0)
    x = torch.randn(3, 50, 64)
    with torch.no_grad():
        X = torch.fft.rfft(x, dim=1, norm='ortho') * torch.view_as_complex(f.complex_weight)
        a = torch.fft.irfft(X, n=50, dim=1, norm='ortho')
        b = orc.fmlp_filter_explicit(x, f.complex_weight.detach())
    assert float((a - b).abs().max()) < 1e-6


def test_adam_explicit_spec_equals_torch():
    torch.manual_seed(1)
    p = torch.nn.Parameter(torch.randn(7, 5))
    opt = torch.optim.Adam([p], lr=1e-3, weight_decay=1e-4)
    pe, m, v = p.detach().clone(), torch.zeros(7, 5), torch.zeros(7, 5)
    for t in range(1, 5):
        g = torch.randn(7, 5)
        p.grad = g.clone()
        opt.step()
        pe, m, v = orc.adam_explicit(pe, g, m, v, t, lr=1e-3, wd=1e-4)
        assert torch.allclose(p.detach(), pe, rtol=0, atol=1e-7)


def test_bce_closed_form():
    torch.manual_seed(2)
    pos = torch.randn(4, 50)
    pos[:, 30:] = -math.inf
    neg = torch.randn(4, 50, 1)
    valid = ~torch.isinf(pos)
    n = valid.sum()
    want = (-(torch.nn.functional.logsigmoid(pos[valid])).sum() + torch.nn.functional.softplus(neg[..., 0][valid]).sum()) / n
    assert abs(float(orc.bce_loss(pos, neg)) - float(want)) < 1e-6
    assert abs(float(orc.bce_loss(pos, neg, reduce=False).sum()) - float(want)) < 1e-6


def test_toys_checkpoint_known_answer():
    """Shipped SASRec checkpoint reproduces its stored val ndcg@20 / recall@20."""
    fx = load_fixture('toys_ckpt.npz')
    N = int(fx['num_items'][''])
    m = load_params(orc.OracleSASRec(N, embed_dim=64), fx['param']).eval()
    hist = fx['val']['in_item_id'].long()
    tgt, slen = fx['val']['item_id'].long(), fx['val']['seqlen'].long()
    dom = fx['domain_items'][''].long().tolist()
    nd, rc = [], []
    for s in range(0, hist.size(0), 4096):
        b = {'in_item_id': hist[s:s + 4096], 'seqlen': slen[s:s + 4096], 'user_hist': hist[s:s + 4096]}
        _, ids = m.topk(b, 100, dom)
        assert torch.equal(ids[:, :20], fx['top20_all_users'][''][s:s + 4096].long())
        hit = orc.hit_matrix(ids, tgt[s:s + 4096])
        nd.append(orc.ndcg_at_k(hit, 20)); rc.append(orc.recall_at_k(hit, 20))
    assert abs(float(torch.cat(nd).mean()) - float(fx['metric']['ndcg@20'])) < 5e-7
    assert abs(float(torch.cat(rc).mean()) - float(fx['metric']['recall@20'])) < 5e-7
