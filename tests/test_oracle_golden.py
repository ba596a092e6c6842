"""CPU tests: the oracle restatement vs the golden vectors dumped from the unmodified reference
(tests/golden/make_golden.py) and vs its own elementary-algebra spec functions."""
import math

import pytest
import torch

from oracle import dr4sr_oracle as orc
from tests.helpers import load_fixture, oracle_from_fixture, rel_err, load_params

CASES = [('sasrec', 'sasrec_d64.npz'), ('sasrec', 'sasrec_d128.npz'), ('gru4rec', 'gru4rec_d64.npz'),
         ('gru4rec', 'gru4rec_d128.npz'), ('gru4rec', 'gru4rec_h256.npz'), ('fmlp', 'fmlp_d64.npz')]


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


@pytest.mark.parametrize('ckpt', ['toys_ckpt.npz', 'beauty_ckpt.npz', 'sport_ckpt.npz', 'yelp_ckpt.npz'])
def test_shipped_checkpoint_known_answer(ckpt):
    """Each shipped SASRec checkpoint reproduces its stored val ndcg@20 / recall@20."""
    fx = load_fixture(ckpt)
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


@pytest.mark.parametrize('name', ['metamodel_sasrec_d64.npz', 'metamodel_sasrec_d128.npz', 'metamodel_fmlp_d64.npz'])
def test_metamodel_inner_step_oracle_matches_reference_golden(name):
    """orc.meta_weighted_loss (+ the oracle SASRec under it) against the loss and sub-model gradients the UNMODIFIED
    reference MetaModel.training_step produced with the same Gumbel noise (model/metamodel.py:169-194)."""
    fx = load_fixture(name)
    o = oracle_from_fixture('fmlp' if 'fmlp' in name else 'sasrec', fx).train()      # (FMLP: the reference's default sub-model)
    batch = fx['batch']
    mm = {k: v.clone().requires_grad_(True) for k, v in fx['meta'].items()}
    per, q = o.training_step(batch, reduce=False, return_query=True)
    loss = orc.meta_weighted_loss(per, q, mm['0.weight'], mm['0.bias'], mm['2.weight'], mm['2.bias'], fx['meta_cfg']['tau'],
                                  float(fx['meta_cfg']['tau_min']), fx['inner']['gumbel'], batch['user_id'], batch['item_id'])
    loss.backward()
    assert abs(float(loss.detach()) - float(fx['inner']['loss'])) <= 1e-6 * abs(float(fx['inner']['loss']))
    for k, p in o.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        assert rel_err(g, fx['inner_grad'][k]) < 1e-5, k
    for k, p in mm.items():
        assert rel_err(p.grad, fx['inner_meta_grad'][k]) < 1e-5, k


def test_bpr_loss_matches_reference_golden():
    """orc.bpr_loss vs the reference's BPRLoss.forward (model/loss_func.py:40-49) on stored random scores."""
    import numpy as np, os
    from tests.helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, 'bpr_loss.npz'))
    got = orc.bpr_loss(torch.from_numpy(z['pos']), torch.from_numpy(z['neg']))
    assert float(got) == float(z['loss'])
