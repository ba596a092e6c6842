"""GPU known-answer tests: the four SASRec checkpoints the reference ships (dataset/*/*/pre-trained_embedding.ckpt,
SURVEY.md section 4) run through the CUDA path -- `load_state_dict` of the reference's own parameter names,
`SASRec.topk` over every validation user -- and must reproduce the reference's top-20 ids and the ndcg@20 /
recall@20 stored inside the checkpoint (reference model/basemodel.py:337-365).  Real data: D = 64, mean sequence
length 2-7, i.e. the shipped configuration and the only non-synthetic input of the suite.

Fixtures: tests/golden/{toys,beauty,sport,yelp}_ckpt.npz, written by tests/golden/make_golden.py from the
unmodified reference (which asserts the reference reproduces the stored metrics before dumping).
"""
import pytest
import torch

from tests.helpers import load_fixture, load_params

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _model(fx, domain='dom'):
    from dr4sr_b200.model.sasrec import SASRec
    from dr4sr_b200.utils.config import SyntheticCatalog, default_config
    N = int(fx['num_items'][''])
    cfg = default_config('SASRec', model__embed_dim=64, train__device=DEV)
    cat = SyntheticCatalog(N, domain=domain, items=fx['domain_items'][''].long().tolist())
    m = SASRec(cfg, [cat] * 3)
    m._init_model()
    load_params(m, {k: v.to(DEV) for k, v in fx['param'].items()})      # reference state_dict keys, unchanged
    m.set_eval_domain(domain)
    return m.eval()


@pytest.mark.parametrize('name', ['toys_ckpt.npz', 'beauty_ckpt.npz', 'sport_ckpt.npz', 'yelp_ckpt.npz'])
def test_shipped_checkpoint_known_answer_on_cuda(name):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    fx = load_fixture(name)
    m = _model(fx)
    hist = fx['val']['in_item_id'].long()
    tgt, slen = fx['val']['item_id'].long(), fx['val']['seqlen'].long()
    want20 = fx['top20_all_users'][''].long()
    want100 = fx['top100_first_users'][''].long()
    nd = rc = 0.0
    bad_rows, worst_gap, smax = 0, 0.0, 0.0
    disc = torch.log2(torch.arange(20, dtype=torch.float64) + 2.0)
    from dr4sr_b200 import engine
    dev_sums = torch.zeros(4, dtype=torch.float64, device=DEV)          # the eval loop's fused metric kernel, cutoffs (20, 10)
    nd10 = rc10 = 0.0
    for s in range(0, hist.size(0), 2048):                     # eval batch_size of the reference (configs/basemodel.yaml)
        e = min(s + 2048, hist.size(0))
        b = {'in_item_id': hist[s:e].to(DEV), 'seqlen': slen[s:e].to(DEV), 'user_hist': hist[s:e].to(DEV)}
        scores, ids = m.topk(b, 100, b['user_hist'])
        ids = ids.cpu()
        assert ids.dtype == torch.int64 and scores.dtype == torch.float32
        assert bool((scores[:, :-1] >= scores[:, 1:]).all()), 'scores not sorted descending'
        # ids must equal the reference's, except where fp32 summation order decides between scores that are equal to within
        # round-off (fp32 vs fp64 already differs in 23 of these 19 412 rows at top-100): a differing
        # position must hold an id whose score here is within a few ulp-of-the-dot-product of the reference id's score here
        sc = scores.cpu()
        smax = max(smax, float(sc[:, 0].abs().max()))
        for r in (ids[:, :20] != want20[s:e]).any(dim=1).nonzero().flatten().tolist():
            bad_rows += 1
            for j in (ids[r, :20] != want20[s + r]).nonzero().flatten().tolist():
                pos = (ids[r] == want20[s + r, j]).nonzero().flatten()
                assert pos.numel() == 1, f'user {s + r}: reference id {int(want20[s + r, j])} (rank {j}) is not in our top-100'
                worst_gap = max(worst_gap, abs(float(sc[r, j]) - float(sc[r, int(pos)])))
        if s == 0:
            n100 = min(want100.size(0), e)
            bad100 = int((ids[:n100] != want100[:n100]).any(dim=1).sum())
        hit = (tgt[s:e].view(-1, 1) == ids[:, :20]).double()
        nd += float((hit / disc).sum())
        rc += float(hit.sum())
        nd10 += float((hit[:, :10] / disc[:10]).sum())
        rc10 += float(hit[:, :10].sum())
        engine.rank_metrics(ids.to(DEV), tgt[s:e].to(DEV), [20, 10], dev_sums)
    n = hist.size(0)
    ndcg, recall = nd / n, rc / n
    got = dev_sums.tolist()                                              # dr4sr_rank_metrics == the torch hit-matrix arithmetic
    for a, b in zip(got, (nd, rc, nd10, rc10)):
        assert abs(a - b) <= 1e-6 * max(1.0, abs(b)), (got, (nd, rc, nd10, rc10))
    print(f'{name}: {n} users, top-20 id rows differing from the reference: {bad_rows}, top-100 rows differing (first {n100}): {bad100}; '
          f'ndcg@20 {ndcg:.7f} (stored {float(fx["metric"]["ndcg@20"]):.7f}), recall@20 {recall:.7f} '
          f'(stored {float(fx["metric"]["recall@20"]):.7f})')
    print(f'   near-tie swaps: {bad_rows} users, largest score gap between swapped ids {worst_gap:.3e} (max |score| {smax:.2f})')
    assert bad_rows <= max(2, n // 2000), f'{bad_rows} users whose top-20 ids differ from the reference'
    assert worst_gap <= 4e-6 * smax, f'ids differ across a score gap of {worst_gap:.3e}: not a round-off tie'
    # the stored metrics to 5e-7, plus at most one rank position per user whose ids differ across a round-off tie
    slack = 5e-7 + bad_rows / n
    assert abs(ndcg - float(fx['metric']['ndcg@20'])) < slack
    assert abs(recall - float(fx['metric']['recall@20'])) < slack


def test_checkpoint_roundtrip_keys_and_save(tmp_path):
    """state_dict keys equal the reference checkpoint's; save_checkpoint writes the reference's dict layout
    (utils/callbacks.py:70-76) and load_checkpoint restores it."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    fx = load_fixture('toys_ckpt.npz')
    m = _model(fx)
    sd = m.state_dict()
    for k, v in fx['param'].items():
        assert k in sd and tuple(sd[k].shape) == tuple(v.shape), k
        assert torch.equal(sd[k].cpu(), v), k
    m.config['eval']['save_path'] = str(tmp_path)
    path = m.save_checkpoint()
    ck = torch.load(path, weights_only=False, map_location='cpu')
    assert set(ck) == {'config', 'model', 'epoch', 'parameters', 'metric'} and ck['model'] == 'SASRec'
    with torch.no_grad():
        m.item_embedding.weight.mul_(0.5)
    m.load_checkpoint(path)
    assert torch.equal(m.state_dict()['item_embedding.weight'].cpu(), fx['param']['item_embedding.weight'])
