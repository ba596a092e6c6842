"""Dataset -> batch feed with the reference's interface (reference data/dataset.py:10-164).

Same files (`dataset/<name>/<domain>/{inter.csv, train<suffix>.pth, val.pth, test.pth}`: python lists of
rows `[uid, in_seq[L], target(s), seqlen, label(s), domain[L], (hist[L])]`), same attributes
(`num_items`, `num_users`, `domain_item_mapping`, ...), same batch dict.  The difference is the loader:
the reference runs `DataLoader(self, batch_size, shuffle)`, i.e. B Python `__getitem__` calls and a
`default_collate` per batch (7.5 ms at B = 256); here the columns live on the device as
int64 matrices and a batch is ONE permutation slice + one gather kernel per column (`dr4sr_gather_i64`).
"""
from __future__ import annotations

import logging
import os
from typing import Dict, Iterator, List, Optional

import torch

from .. import _lib
from ..engine import _p, _stream

TRAIN_KEYS = ('user_id', 'in_item_id', 'item_id', 'seqlen', 'label', 'domain_id')


def _column(rows, k, device) -> torch.Tensor:
    return torch.tensor([r[k] for r in rows], dtype=torch.int64, device=device)


class DeviceBatchLoader:
    """Iterates batches of a dict of device-resident int64 columns (first dimension = rows)."""

    def __init__(self, columns: Dict[str, torch.Tensor], batch_size: int, shuffle: bool, with_index: bool = True) -> None:
        self.columns, self.batch_size, self.shuffle, self.with_index = columns, int(batch_size), shuffle, with_index
        self.n = next(iter(columns.values())).size(0)

    def __len__(self) -> int:
        return (self.n + self.batch_size - 1) // self.batch_size

    def _gather(self, col: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        if not col.is_cuda:
            raise _lib.Dr4srError('the batch feed gathers on the GPU (no CPU fallback): build the dataset with a CUDA device')
        width = 1 if col.dim() == 1 else col.size(1)
        out = torch.empty((idx.numel(),) + tuple(col.shape[1:]), dtype=torch.int64, device=col.device)
        _lib.check(_lib.lib().dr4sr_gather_i64(_p(col), width, _p(idx), idx.numel(), _p(out), _stream()), 'dr4sr_gather_i64')
        return out

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        dev = next(iter(self.columns.values())).device
        order = torch.randperm(self.n, device=dev) if self.shuffle else torch.arange(self.n, device=dev)
        for s in range(0, self.n, self.batch_size):
            idx = order[s:s + self.batch_size].contiguous()
            batch = {k: self._gather(v, idx) for k, v in self.columns.items()}
            if self.with_index:
                batch['index'] = idx
            yield batch


class BaseDataset:
    def __init__(self, config, phase='train') -> None:
        self.name = config['data']['dataset']
        self.fuid, self.fiid = 'user_id', 'item_id'
        self.logger = logging.getLogger('CDR')
        self.config, self.phase = config, phase
        self.device = config['train']['device']
        self.root = config['data'].get('root', 'dataset')
        self.domain_name_list = config['data']['domain_name_list']
        self.max_seq_len = config['data']['max_seq_len']
        self._data = self.data = None
        self._load_datasets()
        self.domain_user_mapping = self._domain_ids('user_id')
        self.domain_item_mapping = self._domain_ids('item_id')
        self.eval_domain = self.domain_name_list[0]

    # ---- counts and domain membership from inter.csv (data/dataset.py:41-65) ------------------------
    def _load_datasets(self) -> None:
        import pandas as pd
        frames = [pd.read_csv(os.path.join(self.root, self.name, d, 'inter.csv')) for d in self.domain_name_list]
        self._inter_data = pd.concat(frames)
        self._num_users = self._inter_data['user_id'].nunique() + 1          # +1: padding id 0
        self._num_items = self._inter_data['item_id'].nunique() + 1

    def _domain_ids(self, column: str) -> Dict[str, List[int]]:
        out = {}
        for i, d in enumerate(self.domain_name_list):
            out[d] = self._inter_data.loc[self._inter_data['domain'] == i, column].unique().tolist()
        return out

    num_users = property(lambda self: self._num_users)
    num_items = property(lambda self: self._num_items)
    num_domains = property(lambda self: len(self.domain_name_list))

    def __len__(self) -> int:
        cols = self.data if self.phase == 'train' else self.data[self.eval_domain]
        return cols['user_id'].size(0)

    # ---- python lists -> device-resident columns (data/dataset.py:79-91) ------------------------------
    def unpack(self, rows) -> Dict[str, torch.Tensor]:
        cols = {k: _column(rows, i, self.device) for i, k in enumerate(TRAIN_KEYS)}
        if self.phase != 'train':
            cols['user_hist'] = cols['in_item_id']
        return cols

    def build(self) -> None:
        self._build()
        self.data = self._data

    def get_loader(self, batch_size: Optional[int] = None, shuffle: bool = True) -> DeviceBatchLoader:
        if self.phase == 'train':
            bs = self.config['train']['batch_size'] if batch_size is None else batch_size
            return DeviceBatchLoader(self.data, bs, shuffle)
        bs = self.config['eval']['batch_size'] if batch_size is None else batch_size
        return DeviceBatchLoader(self.data[self.eval_domain], bs, shuffle=False)

    def set_eval_domain(self, domain) -> None:
        self.eval_domain = domain


class SeparateDataset(BaseDataset):
    """All domains' sequences put together (data/dataset.py:121-164); the class every shipped config uses."""

    def _load_datasets(self) -> None:
        super()._load_datasets()
        suffix = self.config['data']['train_file'] if self.phase == 'train' else ''
        self._raw = [torch.load(os.path.join(self.root, self.name, d, self.phase + suffix + '.pth'), weights_only=False)
                     for d in self.domain_name_list]

    def _build(self) -> None:
        if self.phase == 'train':
            rows = []
            for r in self._raw:
                rows += r
            self._data = self.unpack(rows)
        else:
            self._data = {d: self.unpack(r) for d, r in zip(self.domain_name_list, self._raw)}
        self._raw = None
