"""Container modules with the reference's names (reference module/layers.py:9-136).

On the CUDA path these hold parameters and describe structure; the arithmetic lives in libdr4sr
(pooling is folded into the packed-row layout, the GRU runs in `dr4sr_gru_fwd/bwd`).  They exist so
``state_dict()`` keys (`query_encoder.0.3.gru.weight_ih_l0`, ...) equal the reference's.
"""
from __future__ import annotations

from typing import Tuple

import torch


class SeqPoolingLayer(torch.nn.Module):
    """'origin' (zero rows t >= seqlen) and 'last' (row seqlen-1) pooling of module/layers.py:41-50,69-73,
    served by `dr4sr_unpack_rows` on packed rows."""

    def __init__(self, pooling_type='mean', keepdim=False) -> None:
        super().__init__()
        if pooling_type not in ('origin', 'last'):
            raise NotImplementedError("libdr4sr implements the pooling types the shipped models use: 'origin', 'last'")
        self.pooling_type = pooling_type
        self.keepdim = keepdim


class VStackLayer(torch.nn.Sequential):
    def forward(self, input):
        for module in self:
            input = module(*input) if isinstance(input, Tuple) else module(input)
        return input


class LambdaLayer(torch.nn.Module):
    def __init__(self, lambda_func) -> None:
        super().__init__()
        self.lambda_func = lambda_func

    def forward(self, *args):
        return self.lambda_func(args[0]) if len(args) == 1 else self.lambda_func(args)


class GRULayer(torch.nn.Module):
    """Parameter container for the bias-free multi-layer GRU (module/layers.py:117-136)."""

    def __init__(self, input_dim, output_dim, num_layer=1, bias=False, batch_first=True, bidirectional=False,
                 return_hidden=False) -> None:
        super().__init__()
        if bias or bidirectional or not batch_first:
            raise NotImplementedError('libdr4sr implements the GRU the reference configures: no bias, unidirectional, batch_first')
        self.gru = torch.nn.GRU(input_size=input_dim, hidden_size=output_dim, num_layers=num_layer, bias=bias,
                                batch_first=batch_first, bidirectional=bidirectional)
        self.return_hidden = return_hidden
