"""GPU parity tests of the FMLP path (embedding + LayerNorm, spectral filter as circular convolution,
4D FFN, last-position query, 1-D target BCE) against the golden vectors of the unmodified reference
(D = 64, the only width the reference can run) and against the CPU oracle (D = 128, BASELINE config 3)."""
import pytest
import torch

from oracle import dr4sr_oracle as orc
from tests.helpers import load_fixture, rel_err, load_params

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = {'tc': dict(fwd=1e-4, loss=1e-4, grad=1e-3, adam=1e-4), 'ffma': dict(fwd=1e-5, loss=1e-5, grad=1e-4, adam=2e-5)}


@pytest.fixture(params=['tc', 'ffma'], autouse=True)
def backend(request):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from dr4sr_b200 import _lib
    _lib.check(_lib.lib().dr4sr_set_gemm_backend(0 if request.param == 'tc' else 1), 'set_gemm_backend')
    yield request.param
    _lib.lib().dr4sr_set_gemm_backend(0)


def make_model(N, D, layers=2, p=0.0, seed=2023):
    from dr4sr_b200.model.fmlp import FMLP
    from dr4sr_b200.utils.config import default_config, SyntheticCatalog
    cfg = default_config('FMLP', model__embed_dim=D, model__layer_num=layers, model__dropout_rate=p, train__device=DEV,
                         train__seed=seed)
    torch.manual_seed(seed)
    m = FMLP(cfg, [SyntheticCatalog(N)] * 3)
    m._init_model()
    return m


def to_dev(batch):
    return {k: v.to(DEV) for k, v in batch.items()}


def test_fmlp_matches_reference_golden_d64(backend):
    tol = TOL[backend]
    fx = load_fixture('fmlp_d64.npz')
    N, D = fx['param']['item_embedding.weight'].shape
    m = make_model(N, D).train()
    load_params(m, {k: v.to(DEV) for k, v in fx['param'].items()})
    batch = to_dev(fx['batch'])
    q = m.forward(batch).cpu()
    assert q.shape == fx['train']['query'].shape
    assert rel_err(q, fx['train']['query']) < tol['fwd']
    per = m.training_step(batch, reduce=False).detach().cpu()
    assert rel_err(per, fx['train']['loss_per_pos']) < tol['loss']
    m.optimizer.zero_grad()
    loss = m.training_step(batch)
    loss.backward()
    assert abs(float(loss.detach()) - float(fx['train']['loss'])) / abs(float(fx['train']['loss'])) < tol['loss']
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        assert rel_err(p.grad.cpu(), fx['grad'][k]) < tol['grad'], k
    # irfft drops Im(W[0]) and Im(W[L/2]): their gradients are exactly zero
    g = m.item_encoder.layer[0].filterlayer.complex_weight.grad
    assert float(g[0, 0, :, 1].abs().max()) == 0.0 and float(g[0, 25, :, 1].abs().max()) == 0.0


def test_fmlp_adam_steps_and_eval_match_reference_golden(backend):
    tol = TOL[backend]
    fx = load_fixture('fmlp_d64.npz')
    N, D = fx['param']['item_embedding.weight'].shape
    m = make_model(N, D).train()
    load_params(m, {k: v.to(DEV) for k, v in fx['param'].items()})
    batch = to_dev(fx['batch'])
    for want in fx['adam']['losses'].tolist():
        m.optimizer.zero_grad()
        loss = m.training_step(batch)
        loss.backward()
        m.optimizer.step()
        assert abs(float(loss.detach()) - want) / want < tol['loss']
    for k, p in m.named_parameters():
        assert float((p.detach().cpu() - fx['param_after'][k]).abs().max()) < tol['adam'], k
    load_params(m, {k: v.to(DEV) for k, v in fx['param_after'].items()})
    m.eval()
    ev = to_dev(fx['evalbatch'])
    q = m.forward(ev).cpu()
    assert rel_err(q, fx['eval']['query']) < tol['fwd']
    k = fx['eval']['topk_ids'].shape[1]
    s, i = m.topk(ev, k, ev['user_hist'])
    assert rel_err(s.cpu(), fx['eval']['topk_scores']) < tol['fwd']
    ws = fx['eval']['topk_scores']
    gap = torch.minimum(torch.cat([ws[:, :1] * 0 + 1, (ws[:, :-1] - ws[:, 1:])], 1),
                        torch.cat([(ws[:, :-1] - ws[:, 1:]), ws[:, :1] * 0 + 1], 1))
    firm = gap > 3e-5
    assert torch.equal(i.cpu()[firm], fx['eval']['topk_ids'][firm])


@pytest.mark.parametrize('B,D,N', [(32, 128, 3000), (7, 64, 500)])
def test_fmlp_training_step_matches_oracle(B, D, N, backend):
    from dr4sr_b200.data.synthetic import synthetic_batch
    tol = TOL[backend]
    m = make_model(N, D).train()
    o = orc.OracleFMLP(N, embed_dim=D, dropout_rate=0.0).train()
    o.load_state_dict({k: v.detach().cpu() for k, v in m.state_dict().items()})
    batch = synthetic_batch(B, 50, N, seed=B, layout='pre')
    batch['item_id'][0] = 0                      # one padded target: excluded from the loss and from n
    lo, qo = o.training_step(batch, return_query=True)
    lo.backward()
    loss, q = m.training_step(to_dev(batch), return_query=True)
    loss.backward()
    assert abs(float(loss.detach()) - float(lo.detach())) / abs(float(lo.detach())) < tol['loss']
    assert rel_err(q.detach().cpu(), qo.detach()) < tol['fwd']
    for (k, p), (_, po) in zip(m.named_parameters(), o.named_parameters()):
        want = po.grad if po.grad is not None else torch.zeros_like(po)
        assert rel_err(p.grad.cpu(), want) < tol['grad'], k


def test_fmlp_dropout_backward_consistency(backend):
    from dr4sr_b200.data.synthetic import synthetic_batch
    N, D = 1000, 64
    m = make_model(N, D, p=0.5).train()
    batch = to_dev(synthetic_batch(48, 50, N, seed=4, layout='pre'))
    eng = m.engine
    eng.step = 3
    a = m.forward(batch).clone()
    eng.step = 3
    b_ = m.forward(batch).clone()
    eng.step = 4
    c = m.forward(batch).clone()
    assert torch.equal(a, b_) and not torch.equal(a, c)
    torch.manual_seed(0)
    params = m._flat_parameters()
    direction = [torch.randn_like(p) * (0.02 if p.dim() > 1 else 0.1) for p in params]

    def loss_at(eps):
        for p, d in zip(params, direction):
            p.data.add_(d, alpha=eps)
        eng.step = 20
        l = float(m.training_step(batch).detach())
        for p, d in zip(params, direction):
            p.data.add_(d, alpha=-eps)
        return l

    eng.step = 20
    m.optimizer.zero_grad()
    loss = m.training_step(batch)
    loss.backward()
    analytic = sum(float((p.grad * d).sum()) for p, d in zip(params, direction))
    h = 1e-2
    numeric = (loss_at(h) - loss_at(-h)) / (2 * h)
    assert abs(analytic - numeric) <= 0.03 * max(abs(numeric), 1e-3), (analytic, numeric)
