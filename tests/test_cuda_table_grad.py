"""GPU test of the embedding-gradient scatter-add through the C ABI: one call (dr4sr_table_grad) == the split the training
step uses (dr4sr_table_grad_targets_async on the background stream + input rows + join) == a float64 index_add of the same
rows (reference: nn.Embedding's backward under model/basemodel.py:204-214).  Float atomics: equal up to summation order."""
import pytest
import torch

from tests.helpers import check_err

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('B,L,D,N', [(64, 50, 128, 5000), (7, 9, 64, 300), (1024, 50, 128, 100_000)])
def test_split_scatter_equals_single_call_and_index_add(B, L, D, N):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from dr4sr_b200 import _lib
    from dr4sr_b200.engine import _p, _stream
    lib = _lib.lib()
    g = torch.Generator().manual_seed(B * 31 + L)
    seqlen = torch.randint(1, L + 1, (B,), generator=g)
    tok_off = torch.zeros(B + 1, dtype=torch.int32)
    tok_off[1:] = seqlen.cumsum(0).int()
    T = int(tok_off[-1])
    row_seq = torch.repeat_interleave(torch.arange(B, dtype=torch.int32), seqlen)
    live = torch.arange(L).view(1, -1) < seqlen.view(-1, 1)
    in_ids = torch.randint(1, N, (B, L), generator=g) * live
    item_id = torch.randint(1, N, (B, L), generator=g) * live
    item_id[:, 0] *= (torch.rand(B, generator=g) < 0.8)        # some live slots without a target (id 0): no gradient
    neg = torch.randint(1, N, (B, L), generator=g)
    dx0 = torch.randn(T, D, generator=g)
    q = torch.randn(T, D, generator=g)
    ds = torch.randn(T, 2, generator=g)
    counts = torch.tensor([T, int((item_id != 0).sum())], dtype=torch.int32)
    d = {k: v.to(DEV) for k, v in dict(tok_off=tok_off, row_seq=row_seq, in_ids=in_ids, item_id=item_id, neg=neg, dx0=dx0, q=q, ds=ds,
                                        counts=counts).items()}
    ws = torch.empty(lib.dr4sr_table_grad_workspace_bytes(L, D), dtype=torch.uint8, device=DEV)

    one = torch.zeros(N, D, device=DEV)
    pos1 = torch.zeros(L, D, device=DEV)
    _lib.check(lib.dr4sr_table_grad(_p(d['dx0']), _p(d['q']), _p(d['ds']), _p(d['in_ids']), _p(d['item_id']), _p(d['neg']), _p(d['tok_off']),
                                    _p(d['row_seq']), _p(d['counts']), B, L, D, N, _p(one), _p(pos1), _p(ws), ws.numel(), _stream()), 'table_grad')
    two = torch.zeros(N, D, device=DEV)
    pos2 = torch.zeros(L, D, device=DEV)
    _lib.check(lib.dr4sr_table_grad_targets_async(_p(d['q']), _p(d['ds']), _p(d['item_id']), _p(d['neg']), _p(d['tok_off']), _p(d['row_seq']),
                                                  _p(d['counts']), B, L, D, N, _p(two), _stream()), 'targets_async')
    _lib.check(lib.dr4sr_table_grad(_p(d['dx0']), None, None, _p(d['in_ids']), None, None, _p(d['tok_off']), _p(d['row_seq']), _p(d['counts']),
                                    B, L, D, N, _p(two), _p(pos2), _p(ws), ws.numel(), _stream()), 'table_grad inputs')
    _lib.check(lib.dr4sr_table_grad_targets_join(_stream()), 'join')
    torch.cuda.synchronize()

    # float64 index_add of the same rows
    want = torch.zeros(N, D, dtype=torch.float64)
    flat_in, flat_tg, flat_ng = in_ids[live], item_id[live], neg[live]
    want.index_add_(0, flat_in, dx0.double() * (flat_in != 0).view(-1, 1))
    has = (flat_tg != 0).view(-1, 1)
    want.index_add_(0, flat_tg, q.double() * ds[:, :1].double() * has)
    want.index_add_(0, flat_ng, q.double() * ds[:, 1:].double() * has)
    want[0] = 0
    scale = float(want.abs().max())
    check_err(f'table_grad one call vs f64 index_add (B={B})', float((one.cpu().double() - want).abs().max()) / scale, 2e-6)
    check_err(f'table_grad split vs f64 index_add (B={B})', float((two.cpu().double() - want).abs().max()) / scale, 2e-6)
    assert torch.equal(pos1, pos2)                                # the positional reduction is deterministic
    t_of = (torch.arange(T) - tok_off[row_seq.long()].long())
    pw = torch.zeros(L, D, dtype=torch.float64).index_add_(0, t_of, dx0.double())
    check_err(f'pos_grad vs f64 (B={B})', float((pos1.cpu().double() - pw).abs().max()) / float(pw.abs().max()), 2e-6)
    assert float(one[0].abs().max()) == 0.0 and float(two[0].abs().max()) == 0.0      # padding row untouched


def _scatter_inputs(B, L, D, N, seed, zipf=False):
    g = torch.Generator().manual_seed(seed)
    seqlen = torch.randint(1, L + 1, (B,), generator=g)
    tok_off = torch.zeros(B + 1, dtype=torch.int32)
    tok_off[1:] = seqlen.cumsum(0).int()
    T = int(tok_off[-1])
    row_seq = torch.repeat_interleave(torch.arange(B, dtype=torch.int32), seqlen)
    live = torch.arange(L).view(1, -1) < seqlen.view(-1, 1)

    def ids():
        if zipf:                                                # Zipf(1.0) over 1..N-1: id 1 takes ~1/H(N) of all slots
            w = 1.0 / torch.arange(1, N, dtype=torch.float64)
            return (torch.multinomial(w, B * L, replacement=True, generator=g) + 1).view(B, L)
        return torch.randint(1, N, (B, L), generator=g)
    in_ids, item_id, neg = ids() * live, ids() * live, ids()
    item_id[:, 0] *= (torch.rand(B, generator=g) < 0.8)
    dx0, q, ds = torch.randn(T, D, generator=g), torch.randn(T, D, generator=g), torch.randn(T, 2, generator=g)
    counts = torch.tensor([T, int((item_id != 0).sum())], dtype=torch.int32)
    host = dict(tok_off=tok_off, row_seq=row_seq, in_ids=in_ids, item_id=item_id, neg=neg, dx0=dx0, q=q, ds=ds, counts=counts)
    want = torch.zeros(N, D, dtype=torch.float64)
    fi, ft, fn = in_ids[live], item_id[live], neg[live]
    has = (ft != 0).view(-1, 1)
    want.index_add_(0, fi, dx0.double() * (fi != 0).view(-1, 1))
    want.index_add_(0, ft, q.double() * ds[:, :1].double() * has)
    want.index_add_(0, fn, q.double() * ds[:, 1:].double() * has)
    want[0] = 0
    return host, want


@pytest.mark.parametrize('B,L,D,N,zipf', [(64, 50, 128, 5000, False), (7, 9, 64, 300, False), (1024, 50, 128, 100_000, False),
                                          (1024, 50, 128, 100_000, True), (256, 50, 64, 11_925, True), (3, 50, 128, 17, False)])
def test_sorted_segment_reduction_is_exact_and_bit_reproducible(B, L, D, N, zipf):
    """dr4sr_table_grad_sorted == float64 index_add (tighter than the atomic path on hot rows: fixed-shape sums), twice the
    same bits, and the padding row stays untouched; Zipf ids put thousands of entries on one row (runs crossing many chunks)."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from dr4sr_b200 import _lib
    from dr4sr_b200.engine import _p, _stream
    lib = _lib.lib()
    host, want = _scatter_inputs(B, L, D, N, seed=B * 7 + N, zipf=zipf)
    d = {k: v.to(DEV) for k, v in host.items()}
    ws = torch.empty(lib.dr4sr_table_grad_sorted_workspace_bytes(B, L, D, N), dtype=torch.uint8, device=DEV)
    outs = []
    for _ in range(2):
        tg = torch.zeros(N, D, device=DEV)
        pos = torch.zeros(L, D, device=DEV)
        ws.random_(0, 255)                                      # no dependence on workspace contents
        _lib.check(lib.dr4sr_table_grad_sorted(_p(d['dx0']), _p(d['q']), _p(d['ds']), _p(d['in_ids']), _p(d['item_id']), _p(d['neg']),
                                               _p(d['tok_off']), _p(d['row_seq']), _p(d['counts']), B, L, D, N, _p(tg), _p(pos), _p(ws),
                                               ws.numel(), _stream()), 'table_grad_sorted')
        torch.cuda.synchronize()
        outs.append((tg, pos))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), 'not bit-reproducible'
    scale = float(want.abs().max())
    check_err(f'sorted table_grad vs f64 index_add (B={B}, N={N}, zipf={zipf})', float((outs[0][0].cpu().double() - want).abs().max()) / scale, 1e-6)
    assert float(outs[0][0][0].abs().max()) == 0.0
    # too small a workspace is an error, not a silent overrun
    rc = lib.dr4sr_table_grad_sorted(_p(d['dx0']), _p(d['q']), _p(d['ds']), _p(d['in_ids']), _p(d['item_id']), _p(d['neg']), _p(d['tok_off']),
                                     _p(d['row_seq']), _p(d['counts']), B, L, D, N, _p(outs[0][0]), None, _p(ws), 16, _stream())
    assert rc != 0


def test_deterministic_training_step_repeats_bit_for_bit():
    """config['train']['deterministic_scatter']: two identical SASRec steps give identical table gradients and parameters
    (the default atomic scatter agrees with them to round-off)."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from dr4sr_b200.data.synthetic import synthetic_batch
    from dr4sr_b200.model.sasrec import SASRec
    from dr4sr_b200.utils.config import SyntheticCatalog, default_config
    N, D, B = 3001, 128, 96
    batch = {k: v.to(DEV) for k, v in synthetic_batch(B, 50, N, seed=5).items()}

    def run(det):
        cfg = default_config('SASRec', model__embed_dim=D, model__dropout_rate=0.3, train__device=DEV)
        cfg['train']['deterministic_scatter'] = det
        torch.manual_seed(11)
        m = SASRec(cfg, [SyntheticCatalog(N)] * 3)
        m._init_model()
        m.train()
        for _ in range(2):
            m.optimizer.zero_grad()
            loss = m.training_step(batch)
            loss.backward()
            g = m.item_embedding.weight.grad.clone()
            m.optimizer.step()
        return g, m.item_embedding.weight.data.clone(), float(loss.detach())

    g1, w1, l1 = run(True)
    g2, w2, l2 = run(True)
    assert torch.equal(g1, g2) and torch.equal(w1, w2) and l1 == l2
    g3, w3, _ = run(False)
    check_err('deterministic vs atomic scatter, table gradient', float((g1 - g3).abs().max() / g1.abs().max()), 5e-6)


def test_loss_value_reads_the_same_loss_without_draining_the_stream():
    """BaseModel.loss_value(): the event-gated side-stream read returns exactly float(loss), step after step."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from dr4sr_b200.data.synthetic import synthetic_batch
    from dr4sr_b200.model.sasrec import SASRec
    from dr4sr_b200.utils.config import SyntheticCatalog, default_config
    N, D, B = 2001, 128, 64
    cfg = default_config('SASRec', model__embed_dim=D, train__device=DEV)
    torch.manual_seed(3)
    m = SASRec(cfg, [SyntheticCatalog(N)] * 3)
    m._init_model()
    m.train()
    with pytest.raises(RuntimeError):
        m.loss_value()
    for i in range(4):
        batch = {k: v.to(DEV) for k, v in synthetic_batch(B, 50, N, seed=30 + i).items()}
        m.optimizer.zero_grad()
        loss = m.training_step(batch)
        loss.backward()
        m.optimizer.step()
        early = m.loss_value()
        assert early == float(loss.detach()) and early > 0.0


def test_second_backward_without_step_or_zero_grad_is_refused():
    """The kernels overwrite the gradient buffers; torch would accumulate.  Instead of silently differing, a second backward()
    without optimizer.step() / zero_grad() raises; after zero_grad() the next backward is accepted."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from dr4sr_b200.data.synthetic import synthetic_batch
    from dr4sr_b200.model.sasrec import SASRec
    from dr4sr_b200.utils.config import SyntheticCatalog, default_config
    N, D, B = 501, 64, 8
    cfg = default_config('SASRec', model__embed_dim=D, model__dropout_rate=0.0, train__device=DEV)
    torch.manual_seed(4)
    m = SASRec(cfg, [SyntheticCatalog(N)] * 3)
    m._init_model()
    m.train()
    batch = {k: v.to(DEV) for k, v in synthetic_batch(B, 50, N, seed=9).items()}
    m.optimizer.zero_grad()
    m.training_step(batch).backward()
    g1 = m.item_embedding.weight.grad.clone()
    with pytest.raises(RuntimeError, match='second backward'):
        m.training_step(batch).backward()
    m.optimizer.zero_grad()                                # an explicit discard: the next backward starts from zero again
    m.training_step(batch).backward()
    assert torch.allclose(m.item_embedding.weight.grad, g1, rtol=0, atol=1e-7 * float(g1.abs().max()) + 1e-12)
    m.optimizer.step()
    m.training_step(batch).backward()                      # step() consumed the gradient: accepted
