"""GPU tests of MetaModel (DR4SR+): the weighted inner step (per-slot loss weights + gradient on the query fed
back into the kernels) against the CPU oracle with the same injected Gumbel noise, and the outer
hypergradient step (composite, twice-differentiable) runs and moves only the meta parameters."""
import pytest
import torch

from oracle import dr4sr_oracle as orc
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def make_meta(sub, N, D, p=0.0):
    from dr4sr_b200.model.metamodel import MetaModel
    from dr4sr_b200.utils.config import default_config, SyntheticCatalog
    cfg = default_config('MetaModel', model__embed_dim=D, model__sub_model=sub, model__dropout_rate=p, train__device=DEV,
                         train__batch_size=16)
    torch.manual_seed(11)
    m = MetaModel(cfg, [SyntheticCatalog(N)] * 3)
    m._init_model()
    return m


@pytest.mark.parametrize('sub,layout', [('SASRec', 'post'), ('FMLP', 'pre')])
def test_meta_inner_step_matches_oracle(sub, layout):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from dr4sr_b200 import _lib
    from dr4sr_b200.data.synthetic import synthetic_batch
    _lib.lib().dr4sr_set_gemm_backend(1)                 # exact-fp32 kernels: isolates the weighting logic
    try:
        N, D, B = 400, 64, 12
        m = make_meta(sub, N, D).train()
        batch = synthetic_batch(B, 50, N, seed=21, layout=layout)
        batch['user_id'][::3] = 0                        # pattern rows: weight forced to 1 (metamodel.py:180-183)
        cls = orc.OracleSASRec if sub == 'SASRec' else orc.OracleFMLP
        o = cls(N, embed_dim=D, dropout_rate=0.0).train()
        o.load_state_dict({k: v.detach().cpu() for k, v in m.sub_model.state_dict().items()})
        shape = (B, 50, 2) if sub == 'SASRec' else (B, 2)
        g = -torch.empty(shape).exponential_(generator=torch.Generator().manual_seed(5)).log()
        m._gumbel_override = g.to(DEV)
        mm = [p.detach().cpu() for p in m.meta_module.parameters()]
        per, q = o.training_step(batch, reduce=False, return_query=True)
        want = orc.meta_weighted_loss(per, q, mm[0], mm[1], mm[2], mm[3], m.tau.detach().cpu(), 1.0, g, batch['user_id'], batch['item_id'])
        want.backward()
        m.sub_model.optimizer.zero_grad()
        loss = m.training_step({k: v.to(DEV) for k, v in batch.items()})
        loss.backward()
        assert abs(float(loss.detach()) - float(want.detach())) / abs(float(want.detach())) < 1e-5
        for (k, p), (_, po) in zip(m.sub_model.named_parameters(), o.named_parameters()):
            ref = po.grad if po.grad is not None else torch.zeros_like(po)
            assert rel_err(p.grad.cpu(), ref) < 2e-4, k
    finally:
        _lib.lib().dr4sr_set_gemm_backend(0)


def test_meta_outer_step_runs_and_moves_only_meta_parameters():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from dr4sr_b200.data.synthetic import synthetic_batch
    N, D, B = 300, 64, 16
    m = make_meta('SASRec', N, D).train()

    class OneBatch:
        def __init__(self, seed):
            self.b = synthetic_batch(B, 50, N, seed=seed, with_neg=False)

        def get_loader(self):
            return [self.b]

    m.dataset_list = [OneBatch(1), OneBatch(2), OneBatch(3)]
    m.sub_model.dataset_list = m.dataset_list
    before_meta = [p.detach().clone() for p in m.meta_module.parameters()]
    before_sub = m.sub_model._flat.clone()
    m._outter_loop(nepoch=11)
    moved = sum(float((a - b.detach()).abs().sum()) for a, b in zip(before_meta, m.meta_module.parameters()))
    assert moved > 0.0
    assert torch.equal(before_sub, m.sub_model._flat)
