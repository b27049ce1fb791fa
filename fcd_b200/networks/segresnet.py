"""SegResNet / SegResNetVAE (MONAI 1.5.1 structure; built at get_model.py:145-187) and the reference's
SegResNet_DSA / SegResNetVAE_DSA (networks/segresnet_dsa/segresnet_dsa.py:23-373), forwards on fcd_b200 kernels.

Module tree and state-dict keys equal the reference's: `convInit.conv`, `down_layers.{i}.{j}`, `up_layers`,
`up_samples.{i}.0.conv` / `.1.pixelshuffle.conv_block`, `transformer_layers.{l}.{k}`, `conv_final.2.conv`, `vae_*`.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .blocks import Convolution, TransformerBlock, UpSample, _act_slope, make_norm
from .ms_dsa_net import _require_cuda, _tuple3


def _norm_spec(norm):
    if isinstance(norm, (tuple, list)):
        return (str(norm[0]).lower(), dict(norm[1]))
    return str(norm).lower()


def _check_instance(norm):
    spec = _norm_spec(norm)
    if (spec if isinstance(spec, str) else spec[0]) != "instance":
        raise NotImplementedError("fcd_b200 SegResNet family runs with norm='INSTANCE' (get_model.py:157)")


class ResBlock(nn.Module):
    """MONAI segresnet_block.ResBlock: norm1-act-conv1-norm2-act-conv2 + identity (pre-activation)."""

    def __init__(self, spatial_dims, in_channels, norm, kernel_size=3, act=("RELU", {"inplace": True})):
        super().__init__()
        self.norm1 = make_norm(_norm_spec(norm), in_channels)
        self.norm2 = make_norm(_norm_spec(norm), in_channels)
        self.slope = _act_slope(act)
        self.act = nn.ReLU(inplace=True) if self.slope == 0.0 else nn.LeakyReLU(self.slope, inplace=True)
        self.conv1 = Convolution(in_channels, in_channels, kernel_size, 1, bias=False, padding=1)
        self.conv2 = Convolution(in_channels, in_channels, kernel_size, 1, bias=False, padding=1)

    def forward(self, x):
        y = ops.norm_act(x, mode="instance", slope=self.slope, eps=self.norm1.eps)
        y = ops.conv3d(y, self.conv1.conv.weight, None, 3)
        y = ops.norm_act(y, mode="instance", slope=self.slope, eps=self.norm2.eps)
        y = ops.conv3d(y, self.conv2.conv.weight, None, 3)
        return ops.add(y, x)


class _Seq(nn.Sequential):
    """nn.Sequential whose children are fcd_b200 blocks taking channels-last activations."""

    def forward(self, x):
        for m in self:
            if isinstance(m, nn.Identity):
                continue
            x = m(x)
        return x


class _StridedConv(Convolution):
    """get_conv_layer(stride=2): 3x3x3 stride-2 pad-1 conv, no bias (segresnet_dsa.py:97)."""

    def __init__(self, cin, cout, bias=False):
        super().__init__(cin, cout, 3, 2, bias=bias, padding=1)

    def forward(self, x):
        return ops.conv3d(x, self.conv.weight, self.conv.bias, 3, stride=2, pad=1)


class _UpSampleStage(nn.Sequential):
    """up_samples[i]: 1x1 conv C -> C/2 then UpSample(pixelshuffle) (segresnet_dsa.py:128-145)."""

    def __init__(self, c, mode):
        super().__init__(Convolution(c, c // 2, 1, 1, bias=False, padding=0),
                         UpSample(3, c // 2, c // 2, scale_factor=2, mode=mode, align_corners=False))

    def forward(self, x, skip=None):
        x = ops.conv3d(x, self[0].conv.weight, None, 1)
        return self[1](x, skip, "add" if skip is not None else "plain")


class _FinalConv(nn.Sequential):
    """_make_final_conv: norm, act, 1x1 conv with bias (segresnet_dsa.py:188-193)."""

    def __init__(self, norm, act_mod, slope, cin, cout):
        super().__init__(make_norm(_norm_spec(norm), cin), act_mod, Convolution(cin, cout, 1, 1, bias=True, padding=0))
        self.slope = slope

    def forward(self, x):
        y = ops.norm_act(x, mode="instance", slope=self.slope, eps=self[0].eps)
        return ops.out_conv(y, self[2].conv.weight, self[2].conv.bias)


class SegResNet(nn.Module):
    def __init__(self, spatial_dims=3, init_filters=8, in_channels=1, out_channels=2, dropout_prob=None,
                 act=("RELU", {"inplace": True}), norm=("GROUP", {"num_groups": 8}), norm_name="", num_groups=8,
                 use_conv_final=True, blocks_down=(1, 2, 2, 4), blocks_up=(1, 1, 1), upsample_mode="nontrainable",
                 **dsa_kw):
        super().__init__()
        if spatial_dims != 3:
            raise ValueError("fcd_b200 supports spatial_dims=3")
        _check_instance(norm)
        self.spatial_dims = spatial_dims
        self.init_filters = init_filters
        self.in_channels = in_channels
        self.blocks_down = blocks_down
        self.blocks_up = blocks_up
        self.dropout_prob = dropout_prob
        self.act = act
        self.slope = _act_slope(act)
        self.act_mod = nn.ReLU(inplace=True) if self.slope == 0.0 else nn.LeakyReLU(self.slope, inplace=True)
        self.norm = norm
        self.upsample_mode = str(upsample_mode).lower().split(".")[-1]
        self.use_conv_final = use_conv_final
        self._dsa_kw = dsa_kw
        self.convInit = Convolution(in_channels, init_filters, 3, 1, bias=False, padding=1)
        self.down_layers = self._make_down_layers()
        self.up_layers, self.up_samples = self._make_up_layers()
        if dsa_kw:
            self.patch_embeddings, self.transformer_layers = self._make_transformer_layers()
        self.conv_final = _FinalConv(norm, self.act_mod, self.slope, init_filters, out_channels)
        if dropout_prob is not None:
            self.dropout = nn.Dropout3d(dropout_prob)

    def _make_down_layers(self):
        layers = nn.ModuleList()
        for i, n in enumerate(self.blocks_down):
            c = self.init_filters * 2 ** i
            pre = _StridedConv(c // 2, c) if i > 0 else nn.Identity()
            layers.append(_Seq(pre, *[ResBlock(3, c, norm=self.norm, act=self.act) for _ in range(n)]))
        return layers

    def _make_up_layers(self):
        up_layers, up_samples = nn.ModuleList(), nn.ModuleList()
        n_up = len(self.blocks_up)
        for i in range(n_up):
            c = self.init_filters * 2 ** (n_up - i)
            up_layers.append(_Seq(*[ResBlock(3, c // 2, norm=self.norm, act=self.act)
                                    for _ in range(self.blocks_up[i])]))
            up_samples.append(_UpSampleStage(c, self.upsample_mode))
        return up_layers, up_samples

    def _make_transformer_layers(self):
        kw = self._dsa_kw
        img = _tuple3(kw["dsa_img_size"])
        self.dsa_start_level = kw["dsa_start_level"]
        pes, trs = nn.ModuleList(), nn.ModuleList()
        for i in range(self.dsa_start_level, len(self.blocks_down)):
            c = self.init_filters * 2 ** i
            n_tok = int(np.prod([s // 2 ** i for s in img]))
            pes.append(nn.Identity())
            trs.append(nn.ModuleList([
                TransformerBlock(input_size=n_tok, hidden_size=c, proj_size=kw["dsa_project_size"], num_heads=4,
                                 dropout_rate=kw["dsa_dropout_rate"], pos_embed=kw["dsa_pos_embed"],
                                 sa_type=kw["dsa_sa_type"]) for _ in range(kw["dsa_num_layers"])]))
        return pes, trs

    # -- forward ---------------------------------------------------------------------------------------------
    def encode(self, x):
        ops.prepack_weights(x.device)
        if self.training:
            ops.tick(x.device)
        x = ops.conv3d(x, self.convInit.conv.weight, None, 3)
        if self.dropout_prob is not None:
            x = ops.dropout3d(x, self.dropout.p, self.training)
        down_x = []
        feature = x
        for i, down in enumerate(self.down_layers):
            x = down(x)
            feature = x
            if self._dsa_kw and i >= self.dsa_start_level:
                for blk in self.transformer_layers[i - self.dsa_start_level]:
                    feature = blk(feature)
            down_x.append(feature)
        return feature, down_x

    def _run_up(self, x, down_x):
        for i, (up, upl) in enumerate(zip(self.up_samples, self.up_layers)):
            x = up(x, down_x[i + 1] if down_x is not None else None)
            x = upl(x)
        return x

    def decode(self, x, down_x):
        x = self._run_up(x, down_x)
        if self.use_conv_final:
            return self.conv_final(x)
        return ops.to_ncdhw(x, self.init_filters)

    def forward(self, x):
        _require_cuda(x)
        return self.forward_cl(ops.to_channels_last(x))

    def forward_cl(self, x0):
        x, down_x = self.encode(x0)
        down_x.reverse()
        return self.decode(x, down_x)


class SegResNetVAE(SegResNet):
    def __init__(self, input_image_size, vae_estimate_std=False, vae_default_std=0.3, vae_nz=256, spatial_dims=3,
                 init_filters=8, in_channels=1, out_channels=2, dropout_prob=None, act=("RELU", {"inplace": True}),
                 norm=("GROUP", {"num_groups": 8}), use_conv_final=True, blocks_down=(1, 2, 2, 4), blocks_up=(1, 1, 1),
                 upsample_mode="nontrainable", **dsa_kw):
        super().__init__(spatial_dims=spatial_dims, init_filters=init_filters, in_channels=in_channels,
                         out_channels=out_channels, dropout_prob=dropout_prob, act=act, norm=norm,
                         use_conv_final=use_conv_final, blocks_down=blocks_down, blocks_up=blocks_up,
                         upsample_mode=upsample_mode, **dsa_kw)
        if vae_estimate_std:
            raise NotImplementedError("vae_estimate_std=True is never used by get_model (get_model.py:173)")
        self.input_image_size = _tuple3(input_image_size)
        self.smallest_filters = 16
        zoom = 2 ** (len(self.blocks_down) - 1)
        self.fc_insize = [s // (2 * zoom) for s in self.input_image_size]
        self.vae_estimate_std = vae_estimate_std
        self.vae_default_std = vae_default_std
        self.vae_nz = vae_nz
        v_filters = self.init_filters * zoom
        total = int(self.smallest_filters * np.prod(self.fc_insize))
        self.vae_down = nn.Sequential(make_norm(_norm_spec(norm), v_filters), self.act_mod,
                                      _StridedConv(v_filters, self.smallest_filters, bias=True),
                                      make_norm(_norm_spec(norm), self.smallest_filters), self.act_mod)
        self.vae_fc1 = nn.Linear(total, vae_nz)
        self.vae_fc2 = nn.Linear(total, vae_nz)
        self.vae_fc3 = nn.Linear(vae_nz, total)
        self.vae_fc_up_sample = nn.Sequential(
            Convolution(self.smallest_filters, v_filters, 1, 1, bias=False, padding=0),
            UpSample(3, v_filters, v_filters, scale_factor=2, mode=self.upsample_mode, align_corners=False),
            make_norm(_norm_spec(norm), v_filters), self.act_mod)
        self.vae_conv_final = _FinalConv(norm, self.act_mod, self.slope, init_filters, in_channels)
        self._vae_noise = None

    def set_vae_noise(self, noise):
        """Inject z-noise [B, vae_nz] instead of randn_like (segresnet_dsa.py:332) -- parity tests only."""
        self._vae_noise = noise

    def _get_vae_loss(self, net_input, vae_input):
        """segresnet_dsa.py:322-359 (vae_estimate_std=False)."""
        eps = self.vae_down[0].eps
        v = ops.norm_act(vae_input, mode="instance", slope=self.slope, eps=eps)
        v = self.vae_down[2](v)
        v = ops.norm_act(v, mode="instance", slope=self.slope, eps=eps)
        B, d, h, w, cp = v.shape
        sf, nz = self.smallest_filters, self.vae_nz
        S = d * h * w
        # the reference flattens NCDHW ([c][s]); our activations are [s][c]: permute the Linear's columns / rows
        w1 = self.vae_fc1.weight.view(nz, sf, S).transpose(1, 2)
        if cp != sf:
            w1 = torch.nn.functional.pad(w1, (0, cp - sf))
        w1 = w1.reshape(nz, S * cp)
        z_mean = ops.conv3d(v.reshape(B, 1, 1, 1, S * cp), w1.view(nz, S * cp, 1, 1, 1), self.vae_fc1.bias, k=1)
        z_mean = z_mean.reshape(B, -1)[:, :nz].float()
        noise = self._vae_noise if self._vae_noise is not None else torch.randn_like(z_mean)
        vae_reg_loss = torch.mean(z_mean ** 2)
        z = (z_mean + self.vae_default_std * noise.to(z_mean)).to(torch.bfloat16)
        w3 = self.vae_fc3.weight.view(sf, S, nz).transpose(0, 1)
        b3 = self.vae_fc3.bias.view(sf, S).t()
        if cp != sf:
            w3 = torch.nn.functional.pad(w3, (0, 0, 0, cp - sf))
            b3 = torch.nn.functional.pad(b3, (0, cp - sf))
        nzp = ops.pad16(nz)
        zin = z if nzp == nz else torch.nn.functional.pad(z, (0, nzp - nz))
        v = ops.conv3d(zin.reshape(B, 1, 1, 1, nzp), w3.reshape(S * cp, nz, 1, 1, 1), b3.reshape(S * cp), k=1)
        v = ops.relu_rows(v.reshape(B, d, h, w, cp), self.slope)
        v = ops.conv3d(v, self.vae_fc_up_sample[0].conv.weight, None, 1)
        v = self.vae_fc_up_sample[1](v, None, "plain")
        v = ops.norm_act(v, mode="instance", slope=self.slope, eps=eps)
        v = self._run_up(v, None)
        recon = self.vae_conv_final(v)
        return vae_reg_loss + ops.mse_loss(recon, net_input)

    def forward(self, x):
        _require_cuda(x)
        net_input = x
        f, down_x = self.encode(ops.to_channels_last(x))
        down_x.reverse()
        logits = self.decode(f, down_x)
        if self.training:
            return logits, self._get_vae_loss(net_input, f)
        return logits, None

    def forward_cl(self, x0):
        if self.training:
            raise RuntimeError("forward_cl is the inference entry point (the VAE loss needs the fp32 input)")
        f, down_x = self.encode(x0)
        down_x.reverse()
        return self.decode(f, down_x), None


class SegResNet_DSA(SegResNet):
    """segresnet_dsa.py:23-230."""

    def __init__(self, spatial_dims=3, init_filters=8, in_channels=1, out_channels=2, dropout_prob=None,
                 act=("RELU", {"inplace": True}), norm=("GROUP", {"num_groups": 8}), norm_name="", num_groups=8,
                 use_conv_final=True, blocks_down=(1, 2, 2, 4), blocks_up=(1, 1, 1), upsample_mode="pixelshuffle",
                 interpolate_mode="linear", dsa_img_size=128, dsa_project_size=64, dsa_num_heads=4, dsa_pos_embed=True,
                 dsa_dropout_rate=0.0, dsa_sa_type="parallel", dsa_bias=False, dsa_num_layers=3, dsa_start_level=3):
        super().__init__(spatial_dims=spatial_dims, init_filters=init_filters, in_channels=in_channels,
                         out_channels=out_channels, dropout_prob=dropout_prob, act=act, norm=norm,
                         use_conv_final=use_conv_final, blocks_down=blocks_down, blocks_up=blocks_up,
                         upsample_mode=upsample_mode, dsa_img_size=dsa_img_size, dsa_project_size=dsa_project_size,
                         dsa_pos_embed=dsa_pos_embed, dsa_dropout_rate=dsa_dropout_rate, dsa_sa_type=dsa_sa_type,
                         dsa_num_layers=dsa_num_layers, dsa_start_level=dsa_start_level)


class SegResNetVAE_DSA(SegResNetVAE):
    """segresnet_dsa.py:232-373."""

    def __init__(self, input_image_size, vae_estimate_std=False, vae_default_std=0.3, vae_nz=256, spatial_dims=3,
                 init_filters=8, in_channels=1, out_channels=2, dropout_prob=None, act=("RELU", {"inplace": True}),
                 norm=("GROUP", {"num_groups": 8}), norm_name="", num_groups=8, use_conv_final=True,
                 blocks_down=(1, 2, 2, 4), blocks_up=(1, 1, 1), upsample_mode="pixelshuffle",
                 interpolate_mode="linear", dsa_img_size=128, dsa_project_size=64, dsa_num_heads=4, dsa_pos_embed=True,
                 dsa_dropout_rate=0.0, dsa_sa_type="parallel", dsa_bias=False, dsa_num_layers=3, dsa_start_level=3):
        super().__init__(input_image_size, vae_estimate_std, vae_default_std, vae_nz, spatial_dims, init_filters,
                         in_channels, out_channels, dropout_prob, act, norm, use_conv_final, blocks_down, blocks_up,
                         upsample_mode, dsa_img_size=dsa_img_size, dsa_project_size=dsa_project_size,
                         dsa_pos_embed=dsa_pos_embed, dsa_dropout_rate=dsa_dropout_rate, dsa_sa_type=dsa_sa_type,
                         dsa_num_layers=dsa_num_layers, dsa_start_level=dsa_start_level)
