"""BaseUNet / MS_DSA_NET / MS_DSA_NET_PS with the reference's module tree (networks/ms_dsa_net/ms_dsa_net.py)
and forwards on fcd_b200 kernels.  Inputs/outputs are NCDHW fp32 tensors exactly as the reference modules take
and return them (train.py:374); internally activations are channels-last bf16."""
from __future__ import annotations

from typing import Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .blocks import (Convolution, GeneralUnetrUpBlock, TransformerBlock, UnetrBasicBlock, UnetrUpBlock, apply_norm,
                     make_norm)

_ACT = ("leakyrelu", {"inplace": True, "negative_slope": 0.01})


def _require_cuda(x):
    if not x.is_cuda:
        raise RuntimeError("fcd_b200 models run on CUDA (sm_100a) only; there is no CPU fallback")


def _tuple3(v):
    if isinstance(v, (tuple, list)):
        if len(v) != 3:
            raise ValueError("img_size must have 3 entries")
        return tuple(int(s) for s in v)
    return (int(v),) * 3


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
        return self.forward_cl(ops.to_channels_last(x))

    def forward_cl(self, out):
        """Forward from a channels-last bf16 batch [B,D,H,W,16] (what fcd_sw_gather produces)."""
        ops.prepack_weights(out.device)
        if self.training:
            ops.tick(out.device)
        feats = []
        for i, enc in enumerate(self.encoders):
            out = enc(out)
            if i != self.depth - 1:
                # pool + skip through one node: the two gradients are summed inside the pool's backward kernel
                out, skip = ops.pool_and_skip(out)
                feats.append(skip)
            else:
                feats.append(out)
        for i, dec in enumerate(self.decoders):
            out = dec(out, feats[-(i + 2)])
        return ops.out_conv(out, self.final_conv.weight, self.final_conv.bias)


class _PatchEmbedding(nn.Sequential):
    """patch_embedding{3..6}: get_conv_layer(k=1, conv_only) + GroupNorm(C/2 groups) (ms_dsa_net.py:215-218)."""

    def __init__(self, cin, cout, bias):
        super().__init__(Convolution(cin, cout, 1, 1, bias), make_norm(("group", {"num_groups": cout // 2}), cout))

    def forward(self, x):
        y = ops.conv3d(x, self[0].conv.weight, self[0].conv.bias, k=1)
        return apply_norm(self[1], y, slope=1.0)


class MS_DSA_NET(nn.Module):
    """ms_dsa_net.py:104-407 (and, with pixel-shuffle decoders, MS_DSA_NET_PS 409-726)."""

    _pixelshuffle = False

    def __init__(self, in_channels: int, out_channels: int, img_size: Sequence[int] | int, feature_size: int = 16,
                 project_size: int = 64, num_heads: int = 4, pos_embed: bool = True,
                 norm_name: Union[Tuple, str] = "instance", act_name=_ACT, dropout_rate: float = 0.0, do_ds=True,
                 spatial_dims: int = 3, sa_type="parallel", res_block=True, bias: bool = False, **up_kw) -> None:
        super().__init__()
        self.name = "MS_DSA_NET_PS" if self._pixelshuffle else "MS_DSA_NET"
        self.do_ds = do_ds
        self.num_classes = out_channels
        if not (0 <= dropout_rate <= 1):
            raise AssertionError("dropout_rate should be between 0 and 1.")
        if spatial_dims != 3:
            raise NotImplementedError("fcd_b200 supports spatial_dims=3")
        self.img_size = _tuple3(img_size)
        self.num_layers = 3
        self.proj_size = project_size
        self.upsample_kernel_size = 2
        self.res_block = res_block
        fs = feature_size
        enc_ch = [in_channels, fs, fs * 2, fs * 4, fs * 8, fs * 16, fs * 32]
        for i in range(1, 7):
            setattr(self, f"encoder{i}", UnetrBasicBlock(spatial_dims, enc_ch[i - 1], enc_ch[i], 3, 1, norm_name,
                                                         act_name, self.res_block, bias))
        for lvl, down in ((3, 4), (4, 8), (5, 16), (6, 32)):
            cin, hidden = enc_ch[lvl], enc_ch[lvl] // 2
            setattr(self, f"patch_embedding{lvl}", _PatchEmbedding(cin, hidden, bias))
            n_tok = int(np.prod([s // down for s in self.img_size]))
            setattr(self, f"trans{lvl}", nn.ModuleList([
                TransformerBlock(input_size=n_tok, hidden_size=hidden, proj_size=32 if lvl == 6 else self.proj_size,
                                 num_heads=4, dropout_rate=dropout_rate, pos_embed=pos_embed, sa_type=sa_type)
                for _ in range(self.num_layers)]))
        dec = {5: (fs * 16, fs * 8), 4: (fs * 8, fs * 4), 3: (fs * 4, fs * 2), 2: (fs * 2, fs * 2), 1: (fs * 2, fs)}
        for lvl in (5, 4, 3, 2, 1):
            ci, co = dec[lvl]
            if self._pixelshuffle:
                blk = GeneralUnetrUpBlock(spatial_dims, ci, co, 3, norm_name, act_name, self.res_block, bias,
                                          upsample_mode=up_kw.get("upsample_mode", "pixelshuffle"),
                                          interpolate_mode=up_kw.get("interpolate_mode", "linear"), scale_factor=2)
            else:
                blk = UnetrUpBlock(spatial_dims, ci, co, 3, 2, norm_name, act_name, self.res_block, bias)
            setattr(self, f"decoder{lvl}", blk)
        self.features_dim = fs * 32
        self.out = _UnetOutBlock(fs, out_channels)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def forward(self, x):
        _require_cuda(x)
        if tuple(x.shape[2:]) != self.img_size:
            raise ValueError(f"MS_DSA_NET was built for patches of {self.img_size}, got {tuple(x.shape[2:])} "
                             "(pos_embed / EF are sized by img_size, ms_dsa_net.py:220-233)")
        return self.forward_cl(ops.to_channels_last(x))

    def forward_cl(self, x0):
        """Forward from a channels-last bf16 batch [B,D,H,W,16] (what fcd_sw_gather produces)."""
        ops.prepack_weights(x0.device)
        if self.training:
            ops.tick(x0.device)
        if tuple(x0.shape[1:4]) != self.img_size:
            raise ValueError(f"MS_DSA_NET was built for patches of {self.img_size}, got {tuple(x0.shape[1:4])}")
        # The four transformer stacks (levels 3-6) are independent of each other and of the deeper encoder levels: stack i
        # is forked onto its own stream as soon as x_i exists, while the encoder goes on; each is joined where the decoder
        # needs it.  The deep stacks are chains of tiny latency-bound kernels, so they hide completely behind level 3.
        # Autograd replays every node on its forward stream, so the backward pass overlaps the same way.
        def stack(lvl, xi):
            t = getattr(self, f"patch_embedding{lvl}")(xi)
            for blk in getattr(self, f"trans{lvl}"):
                t = blk(t)
            return t

        ts, br = {}, {}
        # every encoder output has two consumers (the pool to the next level + the decoder skip / the transformer
        # stack): pool_and_skip sums their gradients inside the pool's backward pass
        xp, x1 = ops.pool_and_skip(self.encoder1(x0))
        xp, x2 = ops.pool_and_skip(self.encoder2(xp))
        for lvl in (3, 4, 5, 6):
            x = getattr(self, f"encoder{lvl}")(xp)
            if lvl < 6:
                xp, x = ops.pool_and_skip(x)
                br[lvl] = ops.branch(x.device, key=lvl).hold(x)     # x is rebound below; the side stream still reads it
                with br[lvl]:
                    ts[lvl] = stack(lvl, x)
            else:
                ts[lvl] = stack(lvl, x)
        br[5].join()
        y = self.decoder5(ts[6], ts[5])
        br[4].join()
        y = self.decoder4(y, ts[4])
        br[3].join()
        y = self.decoder3(y, ts[3])
        y = self.decoder2(y, x2)
        y = self.decoder1(y, x1)
        return ops.out_conv(y, self.out.conv.conv.weight, self.out.conv.conv.bias)


class MS_DSA_NET_PS(MS_DSA_NET):
    _pixelshuffle = True

    def __init__(self, *a, upsample_mode="pixelshuffle", interpolate_mode="linear", **k):
        super().__init__(*a, upsample_mode=upsample_mode, interpolate_mode=interpolate_mode, **k)
        self.upsample_mode = str(upsample_mode)
        self.interpolate_mode = str(interpolate_mode)


class _UnetOutBlock(nn.Module):
    """MONAI UnetOutBlock container: keys out.conv.conv.{weight,bias}."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = Convolution(cin, cout, 1, 1, bias=True)
