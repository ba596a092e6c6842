"""Dataset feed: parsing of the reference's file formats (CPU, real amazon-toys files when /root/reference is
present) and the device-side batch loader (GPU)."""
import os

import pytest
import torch

REF = '/root/reference'


def _cfg(root, device, train_file='_regen', bs=256):
    from dr4sr_b200.utils.config import default_config
    cfg = default_config('SASRec', train__device=device, train__batch_size=bs)
    cfg['data'].update({'dataset': 'amazon-toys', 'domain_name_list': ['toy'], 'root': root, 'train_file': train_file})
    return cfg


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'dataset/amazon-toys/toy')), reason='reference datasets not mounted')
def test_parse_real_toys_files_cpu():
    from dr4sr_b200.data.dataset import SeparateDataset
    cfg = _cfg(os.path.join(REF, 'dataset'), 'cpu')
    tr = SeparateDataset(cfg, 'train'); tr.build()
    va = SeparateDataset(cfg, 'val'); va.build()
    assert tr.num_items == 11925
    assert len(tr) == 79046 and tr.data['in_item_id'].shape == (79046, 50)
    assert tr.data['item_id'].shape == (79046, 50) and va.data['toy']['item_id'].dim() == 1
    assert len(va) == 19412 and 'user_hist' in va.data['toy']
    x = tr.data
    lens = (x['in_item_id'] != 0).sum(1)
    assert torch.equal(lens, x['seqlen'])                     # post-padded, seqlen = number of non-pad inputs
    assert 0 not in tr.domain_item_mapping['toy'] and max(tr.domain_item_mapping['toy']) == tr.num_items - 1


def _write_fixture(root, n=300, L=50, N=97):
    import csv
    g = torch.Generator().manual_seed(0)
    d = os.path.join(root, 'amazon-toys', 'toy')
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, 'inter.csv'), 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['user_id', 'item_id', 'domain'])
        for i in range(1, N):
            w.writerow([1 + i % 40, i, 0])
    def rows(train):
        out = []
        for u in range(n):
            ln = int(torch.randint(1, L + 1, (1,), generator=g))
            seq = torch.randint(1, N, (L + 1,), generator=g).tolist()
            ins = seq[:ln] + [0] * (L - ln)
            if train:
                out.append([u % 40, ins, seq[1:ln + 1] + [0] * (L - ln), ln, [1] * L, [0] * L])
            else:
                out.append([u % 40, ins, seq[ln], ln, 1, [0] * L, ins])
        return out
    torch.save(rows(True), os.path.join(d, 'train_regen.pth'))
    torch.save(rows(False), os.path.join(d, 'val.pth'))
    torch.save(rows(False), os.path.join(d, 'test.pth'))


@pytest.mark.gpu
def test_device_batch_loader_matches_indexing(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from dr4sr_b200.data.dataset import SeparateDataset
    _write_fixture(str(tmp_path))
    cfg = _cfg(str(tmp_path), 'cuda:0', bs=64)
    tr = SeparateDataset(cfg, 'train'); tr.build()
    seen = []
    for batch in tr.get_loader():
        idx = batch['index']
        for k in ('user_id', 'in_item_id', 'item_id', 'seqlen', 'label', 'domain_id'):
            assert batch[k].dtype == torch.int64 and torch.equal(batch[k], tr.data[k][idx]), k
        seen.append(idx)
    seen = torch.cat(seen)
    assert seen.numel() == len(tr) and torch.equal(seen.sort().values, torch.arange(len(tr), device='cuda:0'))   # one pass = every row once
    va = SeparateDataset(cfg, 'val'); va.build()
    b = next(iter(va.get_loader()))
    assert b['item_id'].dim() == 1 and torch.equal(b['user_hist'], b['in_item_id'])


@pytest.mark.gpu
def test_fit_one_epoch_through_the_reference_surface(tmp_path):
    """prepare_datasets -> Model(config, datasets) -> fit() -> evaluate(), the call sequence of quickstart/run.py:7-31."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from dr4sr_b200.model.sasrec import SASRec
    _write_fixture(str(tmp_path))
    cfg = _cfg(str(tmp_path), 'cuda:0', bs=64)
    cfg['train']['epochs'] = 2
    cfg['eval']['save_path'] = str(tmp_path / 'saved')
    cfg['eval']['topk'] = 50
    ds_cls = SASRec._get_dataset_class(cfg)
    datasets = [ds_cls(cfg, ph) for ph in ('train', 'val', 'test')]
    for d in datasets:
        d.build()
    torch.manual_seed(0)
    model = SASRec(cfg, datasets)
    model.fit()
    out = model.evaluate()
    assert 'ndcg@20' in out and 0.0 <= out['ndcg@20'] <= 1.0 and os.path.exists(model.ckpt_path)
    assert model.logged_metrics['train_loss_0'] > 0
