"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference); the GPU box uses the committed
``*.npz`` files.  Recipe = SURVEY.md Appendix A: stub the three absent top-level imports
(faiss / matplotlib / torchmetrics -- none is on the hot path), chdir into the reference
(relative ``configs/`` and ``dataset/`` paths), build the reference model classes with either
the real amazon-toys datasets or a 5-attribute fake dataset for synthetic catalogs.

For every case the reference model's state_dict is loaded into the oracle restatement
(``oracle/dr4sr_oracle.py``) and both are run; the script asserts they agree bit-for-bit before
anything is written, so a fixture pins the reference *and* the oracle.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz
"""
from __future__ import annotations

import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'

for _n in ('faiss', 'matplotlib', 'matplotlib.pyplot', 'torchmetrics', 'torchmetrics.functional'):
    sys.modules[_n] = types.ModuleType(_n)
sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
sys.modules['torchmetrics'].functional = sys.modules['torchmetrics.functional']
os.chdir(REF)
sys.path.insert(0, REF)
sys.path.insert(1, REPO)

import numpy as np   # noqa: E402
import torch         # noqa: E402

import utils as ref_utils                      # noqa: E402  (runs set_detect_anomaly(True), utils/utils.py:11)
from utils import load_config, seed_everything  # noqa: E402
import wandb                                    # noqa: E402

from oracle import dr4sr_oracle as orc          # noqa: E402
from dr4sr_b200.data.synthetic import synthetic_batch   # noqa: E402

wandb.init(mode='disabled')
torch.autograd.set_detect_anomaly(False)
torch.set_num_threads(8)


class FakeDataset:
    def __init__(self, num_items: int) -> None:
        self.num_items = num_items
        self.num_users = 1000
        self.domain_name_list = ['syn']
        self.domain_user_mapping = {'syn': [1]}
        self.domain_item_mapping = {'syn': list(range(1, num_items))}


def ref_model(name: str, N: int, D: int, dropout: float = 0.0, seed: int = 2023, hidden: int | None = None):
    cfg = load_config({'model': name, 'dataset': 'amazon-toys'})
    cfg['train']['device'] = 'cpu'
    cfg['model']['embed_dim'] = D
    cfg['model']['dropout_rate'] = dropout
    if hidden is not None:
        cfg['model']['hidden_size'] = hidden
    seed_everything(seed)
    cls = ref_utils.get_model_class(cfg['model'])
    m = cls(cfg, [FakeDataset(N)] * 3)
    m._init_model(None)
    if name == 'FMLP':                      # dropout p is hard-coded 0.5 (fmlp.py:13, layers.py:744,762)
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = dropout
    m.set_eval_domain('syn')
    return m, cfg


def to_np(d):
    return {k: (v.detach().cpu().numpy().copy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


ALIASES = ('query_encoder.item_encoder.weight', 'query_encoder.0.1.weight')   # second names of item_embedding.weight


def pack(prefix, d):
    return {f'{prefix}/{k}': v for k, v in to_np(d).items() if k not in ALIASES}


def run_case(name: str, oracle_cls, okw: dict, N: int, D: int, B: int, layout: str, seed: int, out: str, steps: int = 3):
    L = 50
    ref, cfg = ref_model(name, N, D, hidden=okw.get('hidden_size'))
    o = oracle_cls(N, embed_dim=D, **okw)
    missing = o.load_state_dict(ref.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    batch = synthetic_batch(B, L, N, seed=seed, layout=layout)
    fx = {}
    fx.update(pack('param', ref.state_dict()))
    fx.update(pack('batch', batch))

    # ---- train-mode forward / loss / gradients (dropout 0) ----
    ref.train(); o.train()
    loss_r, q_r = ref.training_step(batch={k: v.clone() for k, v in batch.items()}, reduce=True, return_query=True)
    loss_o, q_o = o.training_step(batch, reduce=True, return_query=True)
    assert torch.equal(q_r, q_o), 'oracle query != reference query'
    assert torch.equal(loss_r, loss_o), 'oracle loss != reference loss'
    per_r = ref.training_step(batch={k: v.clone() for k, v in batch.items()}, reduce=False)
    per_o = o.training_step(batch, reduce=False)
    assert torch.equal(per_r, per_o)
    ref.optimizer.zero_grad()
    loss_r.backward()
    loss_o.backward()
    grads = {}
    for (k, p), (k2, p2) in zip(ref.named_parameters(), o.named_parameters()):
        assert k == k2
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        g2 = p2.grad if p2.grad is not None else torch.zeros_like(p2)
        # CPU index_put_(accumulate=True) adds atomically from several threads: the table gradient's
        # summation order is not reproducible run to run, so gradients are pinned to 1e-5 relative
        assert torch.allclose(g, g2, rtol=1e-5, atol=1e-9), f'grad mismatch {k}: {(g - g2).abs().max()}'
        grads[k] = g
    fx['train/query'] = q_r.detach().numpy()
    fx['train/loss'] = loss_r.detach().numpy()
    fx['train/loss_per_pos'] = per_r.detach().numpy()
    fx.update(pack('grad', grads))

    # ---- `steps` Adam steps on the same batch (dense Adam, basemodel.py:195-199) ----
    opt_o = o.make_adam(lr=cfg['train']['learning_rate'], weight_decay=cfg['train']['weight_decay'])
    o.zero_grad()
    losses = []
    for s in range(steps):
        ref.optimizer.zero_grad()
        l = ref.training_step(batch={k: v.clone() for k, v in batch.items()})
        l.backward(); ref.optimizer.step()
        opt_o.zero_grad()
        l2 = o.training_step(batch)
        l2.backward(); opt_o.step()
        assert torch.allclose(l, l2, rtol=1e-6, atol=0)
        losses.append(float(l.detach()))
    for (k, p), (_, p2) in zip(ref.named_parameters(), o.named_parameters()):
        assert torch.allclose(p, p2, rtol=0, atol=2e-6), f'post-Adam mismatch {k}: {(p - p2).abs().max()}'
    fx['adam/losses'] = np.asarray(losses, dtype=np.float64)
    fx['adam/weight_decay'] = np.asarray(cfg['train']['weight_decay'], dtype=np.float64)
    fx['adam/lr'] = np.asarray(cfg['train']['learning_rate'], dtype=np.float64)
    fx.update(pack('param_after', ref.state_dict()))

    # ---- eval forward + full-catalog top-k (parameters = after the Adam steps) ----
    ref.eval(); o.eval()
    ev = synthetic_batch(B, L, N, seed=seed + 1, layout=layout, eval_mode=True, with_neg=False)
    with torch.no_grad():
        k = min(100, N - 60)
        s_r, i_r = ref.topk(ev, k, ev['user_hist'])
        s_o, i_o = o.topk(ev, k, ref.domain_item_mapping['syn'])
        assert torch.allclose(s_r, s_o, rtol=1e-5, atol=1e-7)
        o.load_state_dict(ref.state_dict())      # re-sync after the tolerance-level Adam drift
        s_o, i_o = o.topk(ev, k, ref.domain_item_mapping['syn'])
        assert torch.equal(i_r, i_o) and torch.equal(s_r, s_o)
        fx['eval/query'] = ref.forward(ev).numpy()
    fx.update(pack('evalbatch', ev))
    fx['eval/topk_scores'] = s_r.numpy()
    fx['eval/topk_ids'] = i_r.numpy()
    np.savez_compressed(os.path.join(HERE, out), **fx)
    print(f'{out}: loss0={losses[0]:.6f} loss{steps - 1}={losses[-1]:.6f} '
          f'({sum(v.nbytes for v in fx.values()) / 1e6:.2f} MB raw)')


def bpr_case(out: str = 'bpr_loss.npz'):
    """BPRLoss.forward of the reference (model/loss_func.py:40-49) on random scores with -inf at pad targets."""
    from model.loss_func import BPRLoss
    torch.manual_seed(5)
    pos = torch.randn(7, 50) * 3
    pos[:, 31:] = -float('inf')
    pos[2, 5:] = -float('inf')
    neg = torch.randn(7, 50, 1) * 3
    want = BPRLoss()(pos.clone(), neg.clone())
    got = orc.bpr_loss(pos, neg)
    assert torch.equal(want, got), (float(want), float(got))
    np.savez_compressed(os.path.join(HERE, out), pos=pos.numpy(), neg=neg.numpy(), loss=want.numpy())
    print(out, 'loss', float(want))


CKPTS = {   # fixture -> (dataset config name, domain / directory name)      SURVEY.md section 4: the four shipped known answers
    'toys_ckpt.npz': ('amazon-toys', 'toy'),
    'beauty_ckpt.npz': ('amazon-beauty', 'beauty'),
    'sport_ckpt.npz': ('amazon-sport', 'sport'),
    'yelp_ckpt.npz': ('yelp', 'yelp'),
}


def checkpoint_case(out: str, dataset: str, domain: str, n_users_topk: int = 256):
    """Known answer: the shipped SASRec checkpoint of `dataset` on its val split."""
    cfg = load_config({'model': 'SASRec', 'dataset': dataset})
    cfg['train']['device'] = 'cpu'
    seed_everything(cfg['train']['seed'])
    ds = ref_utils.prepare_datasets(cfg)
    m = ref_utils.prepare_model(cfg, ds)
    m._init_model(ds[0])
    ck = torch.load(f'dataset/{dataset}/{domain}/pre-trained_embedding.ckpt', weights_only=False, map_location='cpu')
    m.load_state_dict(ck['parameters'])
    m.eval()
    m.set_eval_domain(domain); ds[1].set_eval_domain(domain)
    o = orc.OracleSASRec(ds[0].num_items, embed_dim=64)
    o.load_state_dict(ck['parameters'])
    o.eval()
    uid, hist, tgt, slen, label, dom, _ = ds[1].data[domain]
    batch = {'user_id': uid, 'in_item_id': hist, 'item_id': tgt, 'seqlen': slen, 'user_hist': hist}
    ndcg, rec, ids_all = [], [], []
    with torch.no_grad():
        for s in range(0, len(uid), 2048):
            b = {k: v[s:s + 2048] for k, v in batch.items()}
            sc_r, id_r = m.topk(b, 100, b['user_hist'])
            sc_o, id_o = o.topk(b, 100, m.domain_item_mapping[domain])
            assert torch.equal(id_r, id_o)
            hit = orc.hit_matrix(id_r, b['item_id'])
            ndcg.append(orc.ndcg_at_k(hit, 20)); rec.append(orc.recall_at_k(hit, 20))
            ids_all.append(id_r)
    ndcg, rec, ids_all = torch.cat(ndcg), torch.cat(rec), torch.cat(ids_all)
    stored = ck['metric']
    print(f'{dataset} ckpt: ndcg@20', float(ndcg.mean()), 'recall@20', float(rec.mean()), 'stored', stored)
    assert abs(float(ndcg.mean()) - float(stored['ndcg@20'])) < 5e-7
    assert abs(float(rec.mean()) - float(stored['recall@20'])) < 5e-7
    assert ds[0].num_items < 32767
    dom_items = np.asarray(sorted(m.domain_item_mapping[domain]), dtype=np.int32)
    fx = {f'param/{k}': v.numpy() for k, v in ck['parameters'].items() if k not in ALIASES}
    fx.update({
        'val/in_item_id': hist.numpy().astype(np.int16), 'val/item_id': tgt.numpy().astype(np.int16),
        'val/seqlen': slen.numpy().astype(np.int16), 'domain_items': dom_items,
        'metric/ndcg@20': np.float64(stored['ndcg@20']), 'metric/recall@20': np.float64(stored['recall@20']),
        'top100_first_users': ids_all[:n_users_topk].numpy().astype(np.int16),
        'top20_all_users': ids_all[:, :20].numpy().astype(np.int16),
        'num_items': np.int64(ds[0].num_items),
    })
    np.savez_compressed(os.path.join(HERE, out), **fx)
    print(out, 'written', os.path.getsize(os.path.join(HERE, out)) / 1e6, 'MB')


def metamodel_case(out: str, N: int, D: int, B: int, seed: int, sub: str = 'SASRec'):
    """MetaModel (DR4SR+, sub_model = SASRec): the weighted inner step (model/metamodel.py:169-194) and one outer
    hypergradient step (metamodel.py:123-166 else-branch, utils/utils.py:145-255) of the UNMODIFIED reference, dropout 0.
    The only randomness left is F.gumbel_softmax's noise: it is drawn from a known seed right before the call, and the
    same noise tensor (regenerated here with the same formula, asserted to reproduce the reference's loss) is stored so
    that the CUDA path can inject it."""
    from torch.nn.attention import SDPBackend, sdpa_kernel
    import model.metamodel as ref_meta
    L = 50
    fake = FakeDataset(N)

    class CpuMetaModel(ref_meta.MetaModel):
        # configuration only: the reference hard-codes the sub-model to GPU 0 and opens a DataLoader
        def _register_sub_model(self):
            sc = load_config({'dataset': self.config['data']['dataset'], 'model': self.config['model']['sub_model']})
            sc['train']['device'] = 'cpu'
            sc['model']['embed_dim'] = D
            sc['model']['dropout_rate'] = 0.0
            return ref_utils.get_model_class(sc['model'])(sc, self.dataset_list)

        def current_epoch_metaloaders(self, nepoch):
            return [None]

    cfg = load_config({'model': 'MetaModel', 'dataset': 'amazon-toys'})
    cfg['train']['device'] = 'cpu'
    cfg['model']['sub_model'] = sub
    cfg['model']['embed_dim'] = D
    cfg['model']['dropout_rate'] = 0.0
    seed_everything(seed)
    ref = CpuMetaModel(cfg, [fake] * 3)
    ref._init_model(None)
    for mod in ref.sub_model.modules():              # FMLP hard-codes nn.Dropout(0.5) (model/fmlp.py:13, module/layers.py:740-808)
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    ref.train()
    # make the meta module non-trivial: the reference initialises it N(0, 0.02), which makes every weight ~0.5
    with torch.no_grad():
        for p in ref.meta_module.parameters():
            p.add_(torch.randn_like(p) * 0.3)
    fx = {}
    fx.update(pack('param', ref.sub_model.state_dict()))
    fx.update(pack('meta', ref.meta_module.state_dict()))
    fx['meta_cfg/tau'] = ref.tau.detach().numpy().copy()
    fx['meta_cfg/tau_min'] = np.float64(cfg['model']['tau_min'])
    for k in ('meta_optimizer', 'meta_learning_rate', 'hpo_learning_rate', 'meta_weight_decay'):
        fx[f'meta_cfg/{k}'] = np.asarray(cfg['train'][k])

    def gumbel_like(shape, s):
        torch.manual_seed(s)
        return -torch.empty(shape, dtype=torch.float32).exponential_().log()    # F.gumbel_softmax's first statement

    layout = 'pre' if sub == 'FMLP' else 'post'          # FMLP: pre-padded inputs, one target per sequence (README.md:78)
    gshape = (B, 2) if sub == 'FMLP' else (B, L, 2)
    tr = synthetic_batch(B, L, N, seed=seed, layout=layout)
    tr['user_id'][::3] = 0            # mined patterns keep weight 1 (metamodel.py:180-183)
    va = synthetic_batch(B, L, N, seed=seed + 1, layout=layout)
    fx.update(pack('batch', tr)); fx.update(pack('valbatch', va))
    fx['meta_cfg/sub_model'] = np.asarray(sub)
    g_in = gumbel_like(gshape, 1234)
    fx['inner/gumbel'] = g_in.numpy()

    # ---- inner weighted step ----
    with sdpa_kernel(SDPBackend.MATH):
        ref.sub_model.optimizer.zero_grad()
        torch.manual_seed(1234)
        loss = ref.training_step(batch={k: v.clone() for k, v in tr.items()}, align=False)
        # oracle restatement with the injected noise must agree
        o = (orc.OracleFMLP if sub == 'FMLP' else orc.OracleSASRec)(N, embed_dim=D, dropout_rate=0.0).train()
        o.load_state_dict(ref.sub_model.state_dict())
        per_o, q_o = o.training_step(tr, reduce=False, return_query=True)
        mm = ref.meta_module
        want = orc.meta_weighted_loss(per_o, q_o, mm[0].weight, mm[0].bias, mm[2].weight, mm[2].bias, ref.tau,
                                      cfg['model']['tau_min'], g_in, tr['user_id'], tr['item_id'])
        assert torch.allclose(loss, want, rtol=1e-6, atol=0), (float(loss), float(want))
        loss.backward()
    fx['inner/loss'] = loss.detach().numpy()
    fx.update(pack('inner_grad', {k: (p.grad if p.grad is not None else torch.zeros_like(p))
                                  for k, p in ref.sub_model.named_parameters()}))
    fx.update(pack('inner_meta_grad', {k: (p.grad if p.grad is not None else torch.zeros_like(p))
                                       for k, p in ref.meta_module.named_parameters()}))

    # ---- one outer step: the else-branch of _outter_loop on (va, tr) with known Gumbel noise ----
    g_out = gumbel_like(gshape, 4321)
    fx['outer/gumbel'] = g_out.numpy()
    with sdpa_kernel(SDPBackend.MATH):
        meta_loss = ref.sub_model.training_step(batch={k: v.clone() for k, v in va.items()}, align=False)
        torch.manual_seed(4321)
        meta_train_loss = ref.training_step(batch={k: v.clone() for k, v in tr.items()}, align=False)
        grads = ref.meta_optimizer.step(val_loss=meta_loss, train_loss=meta_train_loss,
                                        aux_params=list(ref.meta_module.parameters()),
                                        parameters=list(ref.sub_model.parameters()), return_grads=True)
    fx['outer/val_loss'] = meta_loss.detach().numpy()
    fx['outer/train_loss'] = meta_train_loss.detach().numpy()
    for (k, _), g in zip(ref.meta_module.named_parameters(), grads):
        fx[f'outer_hypergrad/{k}'] = g.detach().numpy().copy()       # after clip_grad_norm_ (in place on the same tensors)
    fx.update(pack('meta_after', ref.meta_module.state_dict()))
    np.savez_compressed(os.path.join(HERE, out), **fx)
    print(f'{out}: inner loss {float(loss):.6f}, |hypergrad| '
          f'{float(torch.cat([g.reshape(-1) for g in grads]).norm()):.3e} ({sum(v.nbytes for v in fx.values()) / 1e6:.2f} MB raw)')


if __name__ == '__main__':
    run_case('SASRec', orc.OracleSASRec, dict(dropout_rate=0.0), N=300, D=64, B=6, layout='post', seed=11,
             out='sasrec_d64.npz')
    run_case('SASRec', orc.OracleSASRec, dict(dropout_rate=0.0), N=1000, D=128, B=16, layout='post', seed=12,
             out='sasrec_d128.npz')
    run_case('GRU4Rec', orc.OracleGRU4Rec, dict(dropout_rate=0.0, hidden_size=64), N=300, D=64, B=6, layout='post', seed=13,
             out='gru4rec_d64.npz')
    run_case('GRU4Rec', orc.OracleGRU4Rec, dict(dropout_rate=0.0, hidden_size=128), N=600, D=128, B=8, layout='post', seed=14,
             out='gru4rec_d128.npz')
    run_case('FMLP', orc.OracleFMLP, dict(dropout_rate=0.0), N=300, D=64, B=6, layout='pre', seed=15,
             out='fmlp_d64.npz')
    run_case('GRU4Rec', orc.OracleGRU4Rec, dict(dropout_rate=0.0, hidden_size=256), N=300, D=64, B=6, layout='post', seed=16,
             out='gru4rec_h256.npz')       # hidden 256 = configs/gru4rec.yaml; the size the tcgen05 recurrence serves
    bpr_case()
    for _out, (_ds, _dom) in CKPTS.items():
        checkpoint_case(_out, _ds, _dom)
    metamodel_case('metamodel_sasrec_d64.npz', N=300, D=64, B=6, seed=21)
    metamodel_case('metamodel_sasrec_d128.npz', N=1000, D=128, B=16, seed=22)
    metamodel_case('metamodel_fmlp_d64.npz', N=300, D=64, B=12, seed=23, sub='FMLP')    # the reference's default sub-model (configs/metamodel.yaml)
