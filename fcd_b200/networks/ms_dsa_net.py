"""BaseUNet / MS_DSA_NET / MS_DSA_NET_PS with the reference's module tree (networks/ms_dsa_net/ms_dsa_net.py)
and forwards on fcd_b200 kernels.  Inputs/outputs are NCDHW fp32 tensors exactly as the reference modules take
and return them (train.py:374); internally activations are channels-last bf16."""
from __future__ import annotations

from typing import Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .blocks import UnetrBasicBlock, UnetrUpBlock

_ACT = ("leakyrelu", {"inplace": True, "negative_slope": 0.01})


def _require_cuda(x):
    if not x.is_cuda:
        raise RuntimeError("fcd_b200 models run on CUDA (sm_100a) only; there is no CPU fallback")


class BaseUNet(nn.Module):
    """ms_dsa_net.py:20-101."""

    def __init__(self, in_channels: int, out_channels: int, feature_size: int = 16,
                 norm_name: Union[Tuple, str] = "instance", act_name=_ACT, spatial_dims: int = 3,
                 res_block: bool = False, bias: bool = True, depth: int = 5) -> None:
        super().__init__()
        self.name = "BaseUNet"
        self.depth = depth
        self.in_channels, self.out_channels = in_channels, out_channels
        self.encoders = nn.ModuleList()
        ci, co = in_channels, feature_size
        for i in range(depth):
            self.encoders.append(UnetrBasicBlock(spatial_dims, ci, co, 3, 1, norm_name, act_name, res_block, bias))
            if i != depth - 1:
                ci, co = co, co * 2
        self.decoders = nn.ModuleList()
        ci, co = co, co // 2
        for i in range(depth - 1):
            self.decoders.append(UnetrUpBlock(spatial_dims, ci, co, 3, 2, norm_name, act_name, res_block, bias))
            if i != depth - 2:
                ci, co = co, co // 2
        self.final_conv = nn.Conv3d(co, out_channels, kernel_size=1)

    def forward(self, x):
        _require_cuda(x)
        out = ops.to_channels_last(x)
        feats = []
        for i, enc in enumerate(self.encoders):
            out = enc(out)
            feats.append(out)
            if i != self.depth - 1:
                out = ops.max_pool2(out)
        for i, dec in enumerate(self.decoders):
            out = dec(out, feats[-(i + 2)])
        return ops.out_conv(out, self.final_conv.weight, self.final_conv.bias)
