"""Config helpers mirroring the reference's yaml merge (reference utils/utils.py:90-109) for synthetic runs."""
from __future__ import annotations

from typing import Dict, List


def default_config(model: str = 'SASRec', dataset: str = 'synthetic', **over) -> Dict:
    """The merged dict `load_config` produces from configs/{basemodel,<model>,<dataset>}.yaml."""
    cfg = {
        'data': {'dataset': dataset, 'domain_name_list': ['syn'], 'max_seq_len': 50, 'dataset_class': 'general',
                 'train_file': '_regen'},
        'train': {'batch_size': 256, 'early_stop_mode': 'max', 'early_stop_patience': 20, 'epochs': 1000, 'device': 'cuda',
                  'optimizer': 'adam', 'learning_rate': 0.001, 'weight_decay': 0, 'num_neg': 1, 'seed': 2023},
        'model': {'embed_dim': 64, 'loss_fn': 'bce', 'model': model},
        'eval': {'batch_size': 2048, 'cutoff': [20, 10], 'val_metrics': ['ndcg', 'recall'],
                 'test_metrics': ['ndcg', 'recall'], 'topk': 100, 'save_path': './saved/'},
    }
    per_model = {
        'sasrec': {'model': {'hidden_size': 128, 'layer_num': 2, 'head_num': 2, 'dropout_rate': 0.5, 'activation': 'gelu',
                             'layer_norm_eps': 1e-12}},
        'gru4rec': {'model': {'hidden_size': 256, 'dropout_rate': 0.2, 'layer_num': 2},
                    'train': {'learning_rate': 0.001, 'weight_decay': 0.0001}},
        'fmlp': {'model': {'layer_num': 2, 'dropout_rate': 0.5}},
        'metamodel': {'model': {'sub_model': 'FMLP', 'hidden_size': 128, 'layer_num': 2, 'head_num': 2, 'dropout_rate': 0.5,
                                'activation': 'gelu', 'tau_min': 1, 'layer_norm_eps': 1e-12},
                      'train': {'interval': 30, 'meta_optimizer': 'sgd', 'meta_learning_rate': 0.001, 'hpo_learning_rate': 0.001,
                                'meta_weight_decay': 0.001, 'descent_step': 30, 'warmup_epoch': 10, 'early_stop_patience': 20}},
    }[model.lower()]
    for sec, vals in per_model.items():
        cfg[sec].update(vals)
    for key, val in over.items():
        sec, _, leaf = key.partition('__')
        cfg[sec][leaf] = val
    return cfg


class SyntheticCatalog:
    """The five dataset attributes BaseModel.__init__ reads (reference model/basemodel.py:21-42)."""

    def __init__(self, num_items: int, num_users: int = 1000, domain: str = 'syn', items: List[int] | None = None) -> None:
        self.num_items, self.num_users = num_items, num_users
        self.domain_name_list = [domain]
        self.domain_user_mapping = {domain: [1]}
        self.domain_item_mapping = {domain: items if items is not None else range(1, num_items)}
