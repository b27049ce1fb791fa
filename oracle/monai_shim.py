"""TEST INFRASTRUCTURE ONLY -- in-memory stand-in for the `monai==1.5.1` symbols the reference imports.

The reference (/root/reference, mehdirabiee/fcd) delegates its arithmetic to MONAI 1.5.1
(requirements.txt:1), which is NOT installed in this image and is not vendored in the
reference.  To run the reference's OWN network / loss source files on CPU (to validate the
restatement in oracle/nets.py and to generate tests/golden/*), this module registers a
minimal pure-PyTorch `monai` package in sys.modules.  Every class below restates published
MONAI 1.5.1 behaviour from memory ("[RECALLED]" in SURVEY.md Appendix A); parity for those
pieces is therefore UNPINNED against a real MONAI install.  The reference call sites that
need each symbol are cited.

Nothing under fcd_b200/ may import this file.
"""
from __future__ import annotations

import math
import sys
import types
from enum import Enum

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------- monai.utils
class StrEnum(str, Enum):
    def __str__(self):
        return self.value


class InterpolateMode(StrEnum):
    NEAREST = "nearest"
    LINEAR = "linear"
    BILINEAR = "bilinear"
    BICUBIC = "bicubic"
    TRILINEAR = "trilinear"
    AREA = "area"


class UpsampleMode(StrEnum):
    DECONV = "deconv"
    DECONVGROUP = "deconvgroup"
    NONTRAINABLE = "nontrainable"
    PIXELSHUFFLE = "pixelshuffle"


def ensure_tuple_rep(tup, dim):
    """monai.utils.ensure_tuple_rep (ms_dsa_net.py:132, segresnet_dsa.py:56)."""
    if isinstance(tup, torch.Tensor):
        tup = tup.detach().cpu().numpy()
    if isinstance(tup, np.ndarray):
        tup = tup.tolist()
    if not isinstance(tup, (list, tuple)):
        return (tup,) * dim
    if len(tup) == dim:
        return tuple(tup)
    raise ValueError(f"Sequence must have length {dim}, got {len(tup)}.")


# ----------------------------------------------------- monai.networks.layers.factories
class _Factory:
    """Indexable like MONAI's LayerFactory: Conv["conv", 3] -> nn.Conv3d."""

    def __init__(self, table):
        self._t = table
        for k in table:
            setattr(self, k.upper(), k.upper())

    def __getitem__(self, args):
        if isinstance(args, str):
            name, dim = args, None
        else:
            name, *rest = args
            dim = rest[0] if rest else None
        fn = self._t[str(name).lower()]
        return fn(dim) if dim is not None else fn(None)


Conv = _Factory({
    "conv": lambda d: (nn.Conv1d, nn.Conv2d, nn.Conv3d)[d - 1],
    "convtrans": lambda d: (nn.ConvTranspose1d, nn.ConvTranspose2d, nn.ConvTranspose3d)[d - 1],
})
Norm = _Factory({
    "instance": lambda d: (nn.InstanceNorm1d, nn.InstanceNorm2d, nn.InstanceNorm3d)[d - 1],
    "batch": lambda d: (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)[d - 1],
    "group": lambda d: nn.GroupNorm,
    "layer": lambda d: nn.LayerNorm,
})
Act = _Factory({
    "relu": lambda d: nn.ReLU,
    "leakyrelu": lambda d: nn.LeakyReLU,
    "prelu": lambda d: nn.PReLU,
    "gelu": lambda d: nn.GELU,
    "sigmoid": lambda d: nn.Sigmoid,
    "tanh": lambda d: nn.Tanh,
})
Dropout = _Factory({
    "dropout": lambda d: (nn.Dropout, nn.Dropout2d, nn.Dropout3d)[d - 1],
    "alphadropout": lambda d: nn.AlphaDropout,
})
Pool = _Factory({
    "avg": lambda d: (nn.AvgPool1d, nn.AvgPool2d, nn.AvgPool3d)[d - 1],
    "max": lambda d: (nn.MaxPool1d, nn.MaxPool2d, nn.MaxPool3d)[d - 1],
})
Pad = _Factory({
    "constantpad": lambda d: (nn.ConstantPad1d, nn.ConstantPad2d, nn.ConstantPad3d)[d - 1],
})


def _split_args(args):
    if isinstance(args, str):
        return args, {}
    name, kw = args
    return name, dict(kw)


def get_norm_layer(name, spatial_dims=1, channels=1):
    """monai.networks.layers.utils.get_norm_layer (conv_blocks.py:418-419,437; ms_dsa_net.py:217)."""
    if name == "":
        return nn.Identity()
    norm_name, norm_args = _split_args(name)
    norm_type = Norm[norm_name, spatial_dims]
    kw = dict(norm_args)
    lname = str(norm_name).lower()
    if lname in ("instance", "batch"):
        kw.setdefault("num_features", channels)
    elif lname == "group":
        kw.setdefault("num_channels", channels)
    elif lname == "layer":
        kw.setdefault("normalized_shape", channels)
    return norm_type(**kw)


def get_act_layer(name):
    """monai.networks.layers.utils.get_act_layer (conv_blocks.py:417; segresnet_dsa.py:72)."""
    if name == "":
        return nn.Identity()
    act_name, act_args = _split_args(name)
    return Act[act_name](**act_args)


# ------------------------------------------------ monai.networks.blocks.convolutions
def _same_padding(kernel_size, dilation=1):
    k = np.atleast_1d(kernel_size)
    d = np.atleast_1d(dilation)
    p = tuple(int(v) for v in ((k - 1) / 2 * d))
    return p if len(p) > 1 else p[0]


def _stride_minus_kernel_padding(kernel_size, stride):
    k = np.atleast_1d(kernel_size)
    s = np.atleast_1d(stride)
    p = tuple(int(v) for v in (s - k))
    return p if len(p) > 1 else p[0]


class Convolution(nn.Sequential):
    """monai Convolution: child `conv` (+ `adn` only when conv_only=False AND any of act/norm/dropout given)."""

    def __init__(self, spatial_dims, in_channels, out_channels, strides=1, kernel_size=3, adn_ordering="NDA",
                 act="PRELU", norm="INSTANCE", dropout=None, dropout_dim=1, dilation=1, groups=1, bias=True,
                 conv_only=False, is_transposed=False, padding=None, output_padding=None):
        super().__init__()
        if padding is None:
            padding = _same_padding(kernel_size, dilation)
        conv_type = Conv["convtrans" if is_transposed else "conv", spatial_dims]
        if is_transposed:
            if output_padding is None:
                output_padding = _stride_minus_kernel_padding(1, strides)
            conv = conv_type(in_channels, out_channels, kernel_size=kernel_size, stride=strides, padding=padding,
                             output_padding=output_padding, groups=groups, bias=bias, dilation=dilation)
        else:
            conv = conv_type(in_channels, out_channels, kernel_size=kernel_size, stride=strides, padding=padding,
                             dilation=dilation, groups=groups, bias=bias)
        self.add_module("conv", conv)
        if conv_only:
            return
        if act is None and norm is None and dropout is None:
            return
        raise NotImplementedError("ADN is never reached from the reference hot path (SURVEY Appendix A1)")


# ------------------------------------------------ monai.networks.blocks.dynunet_block
def _dyn_get_padding(kernel_size, stride):
    k = np.atleast_1d(kernel_size)
    s = np.atleast_1d(stride)
    p = (k - s + 1) / 2
    if np.min(p) < 0:
        raise AssertionError("padding value should not be negative")
    p = tuple(int(v) for v in p)
    return p if len(p) > 1 else p[0]


def _dyn_get_output_padding(kernel_size, stride, padding):
    k = np.atleast_1d(kernel_size)
    s = np.atleast_1d(stride)
    p = np.atleast_1d(padding)
    op = 2 * p + s - k
    if np.min(op) < 0:
        raise AssertionError("out_padding value should not be negative")
    op = tuple(int(v) for v in op)
    return op if len(op) > 1 else op[0]


def dyn_get_conv_layer(spatial_dims, in_channels, out_channels, kernel_size=3, stride=1, act="PRELU",
                       norm="INSTANCE", dropout=None, bias=False, conv_only=True, is_transposed=False):
    """monai.networks.blocks.dynunet_block.get_conv_layer (conv_blocks.py:393-437, 640-649)."""
    padding = _dyn_get_padding(kernel_size, stride)
    output_padding = None
    if is_transposed:
        output_padding = _dyn_get_output_padding(kernel_size, stride, padding)
    return Convolution(spatial_dims, in_channels, out_channels, strides=stride, kernel_size=kernel_size, act=act,
                       norm=norm, dropout=dropout, bias=bias, conv_only=conv_only, is_transposed=is_transposed,
                       padding=padding, output_padding=output_padding)


class UnetOutBlock(nn.Module):
    """monai UnetOutBlock (ms_dsa_net.py:362): 1x1 conv with bias, keys out.conv.conv.{weight,bias}."""

    def __init__(self, spatial_dims, in_channels, out_channels, dropout=None):
        super().__init__()
        self.conv = dyn_get_conv_layer(spatial_dims, in_channels, out_channels, kernel_size=1, stride=1,
                                       dropout=dropout, bias=True, act=None, norm=None, conv_only=False)

    def forward(self, inp):
        return self.conv(inp)


class MLPBlock(nn.Module):  # import-only (conv_blocks.py:13); used by dead code only
    def __init__(self, hidden_size, mlp_dim, dropout_rate=0.0, **kw):
        super().__init__()
        self.linear1 = nn.Linear(hidden_size, mlp_dim)
        self.linear2 = nn.Linear(mlp_dim, hidden_size)
        self.fn = nn.GELU()
        self.drop1 = nn.Dropout(dropout_rate)
        self.drop2 = nn.Dropout(dropout_rate)

    def forward(self, x):
        return self.drop2(self.linear2(self.drop1(self.fn(self.linear1(x)))))


# ------------------------------------------------ monai.networks.blocks.upsample
def pixelshuffle(x, spatial_dims, scale_factor):
    """monai.networks.utils.pixelshuffle: channel c*f^3 + i*f^2 + j*f + k -> output offset (i,j,k)."""
    dim, factor = spatial_dims, scale_factor
    input_size = list(x.size())
    batch_size, channels = input_size[:2]
    scale_divisor = factor ** dim
    if channels % scale_divisor != 0:
        raise ValueError("channels must be divisible by scale_factor ** spatial_dims")
    org_channels = channels // scale_divisor
    output_size = [batch_size, org_channels] + [d * factor for d in input_size[2:]]
    indices = list(range(2, 2 + 2 * dim))
    indices = indices[dim:] + indices[:dim]
    permute_indices = [0, 1]
    for idx in range(dim):
        permute_indices.extend(indices[idx::dim])
    x = x.reshape([batch_size, org_channels] + [factor] * dim + input_size[2:])
    x = x.permute(permute_indices).reshape(output_size)
    return x


def icnr_init(conv, upsample_factor, init=nn.init.kaiming_normal_):
    out_channels, in_channels, *dims = conv.weight.shape
    scale_factor = upsample_factor ** len(dims)
    oc2 = int(out_channels / scale_factor)
    kernel = torch.zeros([oc2, in_channels] + dims)
    kernel = init(kernel)
    kernel = kernel.transpose(0, 1)
    kernel = kernel.reshape(oc2, in_channels, -1)
    kernel = kernel.repeat(1, 1, scale_factor)
    kernel = kernel.reshape([in_channels, out_channels] + dims)
    kernel = kernel.transpose(0, 1)
    conv.weight.data.copy_(kernel)


class SubpixelUpsample(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels=None, scale_factor=2, conv_block="default",
                 apply_pad_pool=True, bias=True):
        super().__init__()
        if scale_factor <= 0:
            raise ValueError("scale_factor must be positive")
        self.dimensions = spatial_dims
        self.scale_factor = scale_factor
        if conv_block == "default":
            out_channels = out_channels or in_channels
            conv_out_channels = out_channels * (scale_factor ** spatial_dims)
            self.conv_block = Conv["conv", spatial_dims](in_channels, conv_out_channels, kernel_size=3, stride=1,
                                                         padding=1, bias=bias)
            icnr_init(self.conv_block, self.scale_factor)
        elif conv_block is None:
            self.conv_block = nn.Identity()
        else:
            self.conv_block = conv_block
        self.pad_pool = nn.Identity()
        if apply_pad_pool:
            self.pad_pool = nn.Sequential(
                Pad["constantpad", spatial_dims](padding=(self.scale_factor - 1, 0) * spatial_dims, value=0.0),
                Pool["avg", spatial_dims](kernel_size=self.scale_factor, stride=1),
            )

    def forward(self, x):
        x = self.conv_block(x)
        x = pixelshuffle(x, self.dimensions, self.scale_factor)
        x = self.pad_pool(x)
        return x


class UpSample(nn.Sequential):
    """monai UpSample (conv_blocks.py:727-735; segresnet_dsa.py:133-141)."""

    def __init__(self, spatial_dims, in_channels=None, out_channels=None, scale_factor=2, kernel_size=None, size=None,
                 mode=UpsampleMode.DECONV, pre_conv="default", post_conv=None, interp_mode=InterpolateMode.LINEAR,
                 align_corners=True, bias=True, apply_pad_pool=True):
        super().__init__()
        scale_factor_ = ensure_tuple_rep(scale_factor, spatial_dims)
        up_mode = UpsampleMode(mode)
        if up_mode == UpsampleMode.DECONV:
            if not kernel_size:
                kernel_size_ = scale_factor_
                output_padding = padding = 0
            else:
                raise NotImplementedError
            self.add_module("deconv", Conv["convtrans", spatial_dims](
                in_channels=in_channels, out_channels=out_channels or in_channels, kernel_size=kernel_size_,
                stride=scale_factor_, padding=padding, output_padding=output_padding, bias=bias))
        elif up_mode == UpsampleMode.NONTRAINABLE:
            if pre_conv == "default" and (out_channels != in_channels):
                self.add_module("preconv", Conv["conv", spatial_dims](
                    in_channels=in_channels, out_channels=out_channels or in_channels, kernel_size=1, bias=bias))
            interp_mode = InterpolateMode(interp_mode)
            linear_mode = [InterpolateMode.LINEAR, InterpolateMode.BILINEAR, InterpolateMode.TRILINEAR]
            if interp_mode in linear_mode:
                interp_mode = linear_mode[spatial_dims - 1]
            self.add_module("upsample_non_trainable", nn.Upsample(
                size=size, scale_factor=None if size else scale_factor_, mode=interp_mode.value,
                align_corners=align_corners))
        elif up_mode == UpsampleMode.PIXELSHUFFLE:
            self.add_module("pixelshuffle", SubpixelUpsample(
                spatial_dims=spatial_dims, in_channels=in_channels, out_channels=out_channels,
                scale_factor=scale_factor_[0], conv_block=pre_conv, apply_pad_pool=apply_pad_pool, bias=bias))
        else:
            raise NotImplementedError(f"Unsupported upsampling mode {mode}.")


# ------------------------------------------------ monai.networks.blocks.segresnet_block
def seg_get_conv_layer(spatial_dims, in_channels, out_channels, kernel_size=3, stride=1, bias=False):
    return Convolution(spatial_dims, in_channels, out_channels, strides=stride, kernel_size=kernel_size, bias=bias,
                       conv_only=True)


def get_upsample_layer(spatial_dims, in_channels, upsample_mode="nontrainable", scale_factor=2):
    return UpSample(spatial_dims=spatial_dims, in_channels=in_channels, out_channels=in_channels,
                    scale_factor=scale_factor, mode=upsample_mode, interp_mode=InterpolateMode.LINEAR,
                    align_corners=False)


class ResBlock(nn.Module):
    """monai ResBlock: norm1-act-conv1-norm2-act-conv2 + identity (pre-activation)."""

    def __init__(self, spatial_dims, in_channels, norm, kernel_size=3, act=("RELU", {"inplace": True})):
        super().__init__()
        if kernel_size % 2 != 1:
            raise AssertionError("kernel_size should be an odd number.")
        self.norm1 = get_norm_layer(name=norm, spatial_dims=spatial_dims, channels=in_channels)
        self.norm2 = get_norm_layer(name=norm, spatial_dims=spatial_dims, channels=in_channels)
        self.act = get_act_layer(act)
        self.conv1 = seg_get_conv_layer(spatial_dims, in_channels, in_channels, kernel_size=kernel_size)
        self.conv2 = seg_get_conv_layer(spatial_dims, in_channels, in_channels, kernel_size=kernel_size)

    def forward(self, x):
        identity = x
        x = self.norm1(x)
        x = self.act(x)
        x = self.conv1(x)
        x = self.norm2(x)
        x = self.act(x)
        x = self.conv2(x)
        x += identity
        return x


# ------------------------------------------------ monai.networks.nets.segresnet
class SegResNet(nn.Module):
    """monai SegResNet (get_model.py:151-163); structure mirrored by the reference's segresnet_dsa.py:82-230."""

    def __init__(self, spatial_dims=3, init_filters=8, in_channels=1, out_channels=2, dropout_prob=None,
                 act=("RELU", {"inplace": True}), norm=("GROUP", {"num_groups": 8}), norm_name="", num_groups=8,
                 use_conv_final=True, blocks_down=(1, 2, 2, 4), blocks_up=(1, 1, 1),
                 upsample_mode=UpsampleMode.NONTRAINABLE):
        super().__init__()
        self.spatial_dims = spatial_dims
        self.init_filters = init_filters
        self.in_channels = in_channels
        self.blocks_down = blocks_down
        self.blocks_up = blocks_up
        self.dropout_prob = dropout_prob
        self.act = act
        self.act_mod = get_act_layer(act)
        if norm_name:
            norm = ("group", {"num_groups": num_groups})
        self.norm = norm
        self.upsample_mode = UpsampleMode(upsample_mode)
        self.use_conv_final = use_conv_final
        self.convInit = seg_get_conv_layer(spatial_dims, in_channels, init_filters)
        self.down_layers = self._make_down_layers()
        self.up_layers, self.up_samples = self._make_up_layers()
        self.conv_final = self._make_final_conv(out_channels)
        if dropout_prob is not None:
            self.dropout = Dropout["dropout", spatial_dims](dropout_prob)

    def _make_down_layers(self):
        down_layers = nn.ModuleList()
        for i, item in enumerate(self.blocks_down):
            c = self.init_filters * 2 ** i
            pre_conv = seg_get_conv_layer(self.spatial_dims, c // 2, c, stride=2) if i > 0 else nn.Identity()
            down_layers.append(nn.Sequential(
                pre_conv, *[ResBlock(self.spatial_dims, c, norm=self.norm, act=self.act) for _ in range(item)]))
        return down_layers

    def _make_up_layers(self):
        up_layers, up_samples = nn.ModuleList(), nn.ModuleList()
        n_up = len(self.blocks_up)
        for i in range(n_up):
            c = self.init_filters * 2 ** (n_up - i)
            up_layers.append(nn.Sequential(
                *[ResBlock(self.spatial_dims, c // 2, norm=self.norm, act=self.act) for _ in range(self.blocks_up[i])]))
            up_samples.append(nn.Sequential(
                seg_get_conv_layer(self.spatial_dims, c, c // 2, kernel_size=1),
                get_upsample_layer(self.spatial_dims, c // 2, upsample_mode=self.upsample_mode)))
        return up_layers, up_samples

    def _make_final_conv(self, out_channels):
        return nn.Sequential(
            get_norm_layer(name=self.norm, spatial_dims=self.spatial_dims, channels=self.init_filters),
            self.act_mod,
            seg_get_conv_layer(self.spatial_dims, self.init_filters, out_channels, kernel_size=1, bias=True))

    def encode(self, x):
        x = self.convInit(x)
        if self.dropout_prob is not None:
            x = self.dropout(x)
        down_x = []
        for down in self.down_layers:
            x = down(x)
            down_x.append(x)
        return x, down_x

    def decode(self, x, down_x):
        for i, (up, upl) in enumerate(zip(self.up_samples, self.up_layers)):
            x = up(x) + down_x[i + 1]
            x = upl(x)
        if self.use_conv_final:
            x = self.conv_final(x)
        return x

    def forward(self, x):
        x, down_x = self.encode(x)
        down_x.reverse()
        return self.decode(x, down_x)


class SegResNetVAE(SegResNet):
    """monai SegResNetVAE (get_model.py:171-186); VAE branch as mirrored at segresnet_dsa.py:287-373."""

    def __init__(self, input_image_size, vae_estimate_std=False, vae_default_std=0.3, vae_nz=256, spatial_dims=3,
                 init_filters=8, in_channels=1, out_channels=2, dropout_prob=None, act=("RELU", {"inplace": True}),
                 norm=("GROUP", {"num_groups": 8}), use_conv_final=True, blocks_down=(1, 2, 2, 4),
                 blocks_up=(1, 1, 1), upsample_mode=UpsampleMode.NONTRAINABLE):
        super().__init__(spatial_dims=spatial_dims, init_filters=init_filters, in_channels=in_channels,
                         out_channels=out_channels, dropout_prob=dropout_prob, act=act, norm=norm,
                         use_conv_final=use_conv_final, blocks_down=blocks_down, blocks_up=blocks_up,
                         upsample_mode=upsample_mode)
        self.input_image_size = ensure_tuple_rep(input_image_size, spatial_dims)
        self.smallest_filters = 16
        zoom = 2 ** (len(self.blocks_down) - 1)
        self.fc_insize = [s // (2 * zoom) for s in self.input_image_size]
        self.vae_estimate_std = vae_estimate_std
        self.vae_default_std = vae_default_std
        self.vae_nz = vae_nz
        self._prepare_vae_modules()
        self.vae_conv_final = self._make_final_conv(in_channels)

    def _prepare_vae_modules(self):
        zoom = 2 ** (len(self.blocks_down) - 1)
        v_filters = self.init_filters * zoom
        total_elements = int(self.smallest_filters * np.prod(self.fc_insize))
        self.vae_down = nn.Sequential(
            get_norm_layer(name=self.norm, spatial_dims=self.spatial_dims, channels=v_filters),
            self.act_mod,
            seg_get_conv_layer(self.spatial_dims, v_filters, self.smallest_filters, stride=2, bias=True),
            get_norm_layer(name=self.norm, spatial_dims=self.spatial_dims, channels=self.smallest_filters),
            self.act_mod)
        self.vae_fc1 = nn.Linear(total_elements, self.vae_nz)
        self.vae_fc2 = nn.Linear(total_elements, self.vae_nz)
        self.vae_fc3 = nn.Linear(self.vae_nz, total_elements)
        self.vae_fc_up_sample = nn.Sequential(
            seg_get_conv_layer(self.spatial_dims, self.smallest_filters, v_filters, kernel_size=1),
            get_upsample_layer(self.spatial_dims, v_filters, upsample_mode=self.upsample_mode),
            get_norm_layer(name=self.norm, spatial_dims=self.spatial_dims, channels=v_filters),
            self.act_mod)

    def _get_vae_loss(self, net_input, vae_input):
        x_vae = self.vae_down(vae_input)
        x_vae = x_vae.view(-1, self.vae_fc1.in_features)
        z_mean = self.vae_fc1(x_vae)
        z_mean_rand = torch.randn_like(z_mean)
        z_mean_rand.requires_grad_(False)
        if self.vae_estimate_std:
            z_sigma = F.softplus(self.vae_fc2(x_vae))
            vae_reg_loss = 0.5 * torch.mean(z_mean ** 2 + z_sigma ** 2 - torch.log(1e-8 + z_sigma ** 2) - 1)
            x_vae = z_mean + z_sigma * z_mean_rand
        else:
            z_sigma = self.vae_default_std
            vae_reg_loss = torch.mean(z_mean ** 2)
            x_vae = z_mean + z_sigma * z_mean_rand
        x_vae = self.vae_fc3(x_vae)
        x_vae = self.act_mod(x_vae)
        x_vae = x_vae.view([-1, self.smallest_filters] + self.fc_insize)
        x_vae = self.vae_fc_up_sample(x_vae)
        for up, upl in zip(self.up_samples, self.up_layers):
            x_vae = up(x_vae)
            x_vae = upl(x_vae)
        x_vae = self.vae_conv_final(x_vae)
        vae_mse_loss = F.mse_loss(net_input, x_vae)
        return vae_reg_loss + vae_mse_loss

    def forward(self, x):
        net_input = x
        x, down_x = self.encode(x)
        down_x.reverse()
        vae_input = x
        x = self.decode(x, down_x)
        if self.training:
            return x, self._get_vae_loss(net_input, vae_input)
        return x, None


# ------------------------------------------------ monai.losses
def one_hot(labels, num_classes, dtype=torch.float, dim=1):
    if labels.ndim < dim + 1:
        shape = list(labels.shape) + [1] * (dim + 1 - len(labels.shape))
        labels = torch.reshape(labels, shape)
    sh = list(labels.shape)
    if sh[dim] != 1:
        raise AssertionError("labels should have a channel with length equal to one.")
    sh[dim] = num_classes
    o = torch.zeros(size=sh, dtype=dtype, device=labels.device)
    return o.scatter_(dim=dim, index=labels.long(), value=1)


class DiceLoss(nn.Module):
    def __init__(self, include_background=True, to_onehot_y=False, sigmoid=False, softmax=False, other_act=None,
                 squared_pred=False, jaccard=False, reduction="mean", smooth_nr=1e-5, smooth_dr=1e-5, batch=False,
                 weight=None, soft_label=False):
        super().__init__()
        self.include_background, self.to_onehot_y = include_background, to_onehot_y
        self.sigmoid, self.softmax = sigmoid, softmax
        self.squared_pred, self.jaccard = squared_pred, jaccard
        self.reduction = reduction
        self.smooth_nr, self.smooth_dr = float(smooth_nr), float(smooth_dr)
        self.batch = batch
        weight = torch.as_tensor(weight) if weight is not None else None
        self.register_buffer("class_weight", weight)

    def forward(self, input, target):
        if self.sigmoid:
            input = torch.sigmoid(input)
        n_pred_ch = input.shape[1]
        if self.softmax and n_pred_ch != 1:
            input = torch.softmax(input, 1)
        if self.to_onehot_y and n_pred_ch != 1:
            target = one_hot(target, num_classes=n_pred_ch)
        if not self.include_background and n_pred_ch != 1:
            target = target[:, 1:]
            input = input[:, 1:]
        if target.shape != input.shape:
            raise AssertionError(f"ground truth has different shape ({target.shape}) from input ({input.shape})")
        reduce_axis = torch.arange(2, len(input.shape)).tolist()
        if self.batch:
            reduce_axis = [0] + reduce_axis
        intersection = torch.sum(target * input, dim=reduce_axis)
        if self.squared_pred:
            ground_o = torch.sum(target ** 2, dim=reduce_axis)
            pred_o = torch.sum(input ** 2, dim=reduce_axis)
        else:
            ground_o = torch.sum(target, dim=reduce_axis)
            pred_o = torch.sum(input, dim=reduce_axis)
        denominator = ground_o + pred_o
        if self.jaccard:
            denominator = 2.0 * (denominator - intersection)
        f = 1.0 - (2.0 * intersection + self.smooth_nr) / (denominator + self.smooth_dr)
        num_of_classes = target.shape[1]
        if self.class_weight is not None and num_of_classes != 1:
            cw = self.class_weight
            if cw.ndim == 0:
                cw = torch.as_tensor([cw] * num_of_classes)
            f = f * cw.to(f)
        if self.reduction == "mean":
            f = torch.mean(f)
        elif self.reduction == "sum":
            f = torch.sum(f)
        return f


class DiceCELoss(nn.Module):
    def __init__(self, include_background=True, to_onehot_y=False, sigmoid=False, softmax=False, other_act=None,
                 squared_pred=False, jaccard=False, reduction="mean", smooth_nr=1e-5, smooth_dr=1e-5, batch=False,
                 weight=None, lambda_dice=1.0, lambda_ce=1.0, label_smoothing=0.0):
        super().__init__()
        dice_weight = weight[1:] if (weight is not None and not include_background) else weight
        self.dice = DiceLoss(include_background=include_background, to_onehot_y=to_onehot_y, sigmoid=sigmoid,
                             softmax=softmax, squared_pred=squared_pred, jaccard=jaccard, reduction=reduction,
                             smooth_nr=smooth_nr, smooth_dr=smooth_dr, batch=batch, weight=dice_weight)
        self.cross_entropy = nn.CrossEntropyLoss(weight=weight, reduction=reduction, label_smoothing=label_smoothing)
        self.binary_cross_entropy = nn.BCEWithLogitsLoss(pos_weight=weight, reduction=reduction)
        self.lambda_dice, self.lambda_ce = lambda_dice, lambda_ce

    def ce(self, input, target):
        n_pred_ch, n_target_ch = input.shape[1], target.shape[1]
        if n_pred_ch != n_target_ch and n_target_ch == 1:
            target = torch.squeeze(target, dim=1).long()
        elif not torch.is_floating_point(target):
            target = target.to(dtype=input.dtype)
        return self.cross_entropy(input, target)

    def forward(self, input, target):
        dice_loss = self.dice(input, target)
        ce_loss = self.ce(input, target) if input.shape[1] != 1 else self.binary_cross_entropy(input, target.float())
        return self.lambda_dice * dice_loss + self.lambda_ce * ce_loss


def sigmoid_focal_loss(input, target, gamma=2.0, alpha=None):
    loss = input - input * target - F.logsigmoid(input)
    invprobs = F.logsigmoid(-input * (target * 2 - 1))
    loss = (invprobs * gamma).exp() * loss
    if alpha is not None:
        loss = (target * alpha + (1 - target) * (1 - alpha)) * loss
    return loss


class FocalLoss(nn.Module):
    def __init__(self, include_background=True, to_onehot_y=False, gamma=2.0, alpha=None, weight=None,
                 reduction="mean", use_softmax=False):
        super().__init__()
        self.include_background, self.to_onehot_y = include_background, to_onehot_y
        self.gamma, self.alpha, self.reduction, self.use_softmax = gamma, alpha, reduction, use_softmax
        weight = torch.as_tensor(weight) if weight is not None else None
        self.register_buffer("class_weight", weight)

    def forward(self, input, target):
        n_pred_ch = input.shape[1]
        if self.to_onehot_y and n_pred_ch != 1:
            target = one_hot(target, num_classes=n_pred_ch)
        if not self.include_background and n_pred_ch != 1:
            target = target[:, 1:]
            input = input[:, 1:]
        input = input.float()
        target = target.float()
        if self.use_softmax:
            raise NotImplementedError
        loss = sigmoid_focal_loss(input, target, self.gamma, self.alpha)
        num_of_classes = target.shape[1]
        if self.class_weight is not None and num_of_classes != 1:
            raise NotImplementedError
        if self.reduction == "mean":
            loss = loss.mean(dim=list(range(2, len(target.shape))))
            loss = loss.mean()
        elif self.reduction == "sum":
            loss = loss.mean(dim=list(range(2, len(target.shape)))).sum()
        return loss


class DiceFocalLoss(nn.Module):
    def __init__(self, include_background=True, to_onehot_y=False, sigmoid=False, softmax=False, other_act=None,
                 squared_pred=False, jaccard=False, reduction="mean", smooth_nr=1e-5, smooth_dr=1e-5, batch=False,
                 gamma=2.0, focal_weight=None, weight=None, lambda_dice=1.0, lambda_focal=1.0, alpha=None):
        super().__init__()
        weight = focal_weight if focal_weight is not None else weight
        self.dice = DiceLoss(include_background=include_background, to_onehot_y=False, sigmoid=sigmoid,
                             softmax=softmax, squared_pred=squared_pred, jaccard=jaccard, reduction=reduction,
                             smooth_nr=smooth_nr, smooth_dr=smooth_dr, batch=batch, weight=weight)
        self.focal = FocalLoss(include_background=include_background, to_onehot_y=False, gamma=gamma, weight=weight,
                               alpha=alpha, reduction=reduction)
        self.lambda_dice, self.lambda_focal = lambda_dice, lambda_focal
        self.to_onehot_y = to_onehot_y

    def forward(self, input, target):
        if self.to_onehot_y and input.shape[1] != 1:
            target = one_hot(target, num_classes=input.shape[1])
        return self.lambda_dice * self.dice(input, target) + self.lambda_focal * self.focal(input, target)


class GeneralizedDiceLoss(nn.Module):
    """MONAI 1.5.1 monai/losses/dice.py GeneralizedDiceLoss (Sudre et al. 2017), restated from its published
    algorithm: per-class weights 1/G^2 | 1/G | 1 of the label volumes, infinite weights (absent class) replaced by
    the largest finite one, ONE ratio over the weighted class sums."""

    def __init__(self, include_background=True, to_onehot_y=False, sigmoid=False, softmax=False, other_act=None,
                 w_type="square", reduction="mean", smooth_nr=1e-5, smooth_dr=1e-5, batch=False, soft_label=False):
        super().__init__()
        if str(w_type) not in ("square", "simple", "uniform"):
            raise ValueError(f"w_type {w_type!r}")
        self.include_background, self.to_onehot_y = include_background, to_onehot_y
        self.sigmoid, self.softmax, self.w_type, self.reduction = sigmoid, softmax, str(w_type), reduction
        self.smooth_nr, self.smooth_dr, self.batch = float(smooth_nr), float(smooth_dr), batch

    def w_func(self, grnd):
        if self.w_type == "simple":
            return torch.reciprocal(grnd)
        if self.w_type == "square":
            return torch.reciprocal(grnd * grnd)
        return torch.ones_like(grnd)

    def forward(self, input, target):
        if self.sigmoid:
            input = torch.sigmoid(input)
        n_pred_ch = input.shape[1]
        if self.softmax and n_pred_ch != 1:
            input = torch.softmax(input, 1)
        if self.to_onehot_y and n_pred_ch != 1:
            target = one_hot(target, num_classes=n_pred_ch)
        if not self.include_background and n_pred_ch != 1:
            target = target[:, 1:]
            input = input[:, 1:]
        if target.shape != input.shape:
            raise AssertionError(f"ground truth has differing shape ({target.shape}) from input ({input.shape})")
        reduce_axis = torch.arange(2, len(input.shape)).tolist()
        if self.batch:
            reduce_axis = [0] + reduce_axis
        intersection = torch.sum(target * input, reduce_axis)
        ground_o = torch.sum(target, reduce_axis)
        pred_o = torch.sum(input, reduce_axis)
        denominator = ground_o + pred_o
        w = self.w_func(ground_o.float())
        infs = torch.isinf(w)
        if self.batch:
            w[infs] = 0.0
            w = w + infs * torch.max(w)
        else:
            w[infs] = 0.0
            max_values = torch.max(w, dim=1)[0].unsqueeze(dim=1)
            w = w + infs * max_values
        final_reduce_dim = 0 if self.batch else 1
        numer = 2.0 * (intersection * w).sum(final_reduce_dim, keepdim=True) + self.smooth_nr
        denom = (denominator * w).sum(final_reduce_dim, keepdim=True) + self.smooth_dr
        f = 1.0 - (numer / denom)
        if self.reduction == "mean":
            f = torch.mean(f)
        elif self.reduction == "sum":
            f = torch.sum(f)
        return f


class GeneralizedDiceFocalLoss(nn.Module):
    """MONAI 1.5.1 GeneralizedDiceFocalLoss: lambda_gdl * GeneralizedDiceLoss + lambda_focal * FocalLoss, both with the
    caller's include_background / to_onehot_y."""

    def __init__(self, include_background=True, to_onehot_y=False, sigmoid=False, softmax=False, other_act=None,
                 w_type="square", reduction="mean", smooth_nr=1e-5, smooth_dr=1e-5, batch=False, gamma=2.0,
                 focal_weight=None, weight=None, lambda_gdl=1.0, lambda_focal=1.0):
        super().__init__()
        self.generalized_dice = GeneralizedDiceLoss(include_background=include_background, to_onehot_y=to_onehot_y,
                                                    sigmoid=sigmoid, softmax=softmax, w_type=w_type,
                                                    reduction=reduction, smooth_nr=smooth_nr, smooth_dr=smooth_dr,
                                                    batch=batch)
        weight = focal_weight if focal_weight is not None else weight
        self.focal = FocalLoss(include_background=include_background, to_onehot_y=to_onehot_y, gamma=gamma,
                               weight=weight, reduction=reduction)
        self.lambda_gdl, self.lambda_focal = lambda_gdl, lambda_focal

    def forward(self, input, target):
        return self.lambda_gdl * self.generalized_dice(input, target) + self.lambda_focal * self.focal(input, target)


class _Unavailable(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("out of scope (SURVEY.md section 2): not restated by the shim")


# ------------------------------------------------ monai.inferers
def _get_scan_interval(image_size, roi_size, num_spatial_dims, overlap):
    scan_interval = []
    for i in range(num_spatial_dims):
        if roi_size[i] == image_size[i]:
            scan_interval.append(int(roi_size[i]))
        else:
            interval = int(roi_size[i] * (1 - overlap[i]))
            scan_interval.append(interval if interval > 0 else 1)
    return tuple(scan_interval)


def dense_patch_slices(image_size, patch_size, scan_interval):
    num_spatial_dims = len(image_size)
    scan_num = []
    for i in range(num_spatial_dims):
        if scan_interval[i] == 0:
            scan_num.append(1)
        else:
            num = int(math.ceil(float(image_size[i]) / scan_interval[i]))
            scan_dim = next((d for d in range(num) if d * scan_interval[i] + patch_size[i] >= image_size[i]), None)
            scan_num.append(scan_dim + 1 if scan_dim is not None else 1)
    starts = []
    for dim in range(num_spatial_dims):
        dim_starts = []
        for idx in range(scan_num[dim]):
            start_idx = idx * scan_interval[dim]
            start_idx -= max(start_idx + patch_size[dim] - image_size[dim], 0)
            dim_starts.append(start_idx)
        starts.append(dim_starts)
    out = np.asarray([x.flatten() for x in np.meshgrid(*starts, indexing="ij")]).T
    return [tuple(slice(int(s), int(s) + patch_size[d]) for d, s in enumerate(x)) for x in out]


def sliding_window_inference(inputs, roi_size, sw_batch_size, predictor, overlap=0.25, mode="constant",
                             sigma_scale=0.125, padding_mode="constant", cval=0.0, **kwargs):
    """monai.inferers.sliding_window_inference, mode='constant' path (train.py:156-162)."""
    if str(mode) != "constant":
        raise NotImplementedError
    compute_dtype = inputs.dtype
    num_spatial_dims = len(inputs.shape) - 2
    overlap = ensure_tuple_rep(overlap, num_spatial_dims)
    batch_size, _, *image_size_ = inputs.shape
    roi_size = ensure_tuple_rep(roi_size, num_spatial_dims)
    roi_size = tuple(int(r) if r and r > 0 else int(s) for r, s in zip(roi_size, image_size_))
    image_size = tuple(max(image_size_[i], roi_size[i]) for i in range(num_spatial_dims))
    pad_size = []
    for k in range(len(inputs.shape) - 1, 1, -1):
        diff = max(roi_size[k - 2] - inputs.shape[k], 0)
        half = diff // 2
        pad_size.extend([half, diff - half])
    if any(pad_size):
        inputs = F.pad(inputs, pad=pad_size, mode=padding_mode, value=cval)
    scan_interval = _get_scan_interval(image_size, roi_size, num_spatial_dims, overlap)
    slices = dense_patch_slices(image_size, roi_size, scan_interval)
    num_win = len(slices)
    total_slices = num_win * batch_size
    importance_map = torch.ones(roi_size, device=inputs.device, dtype=compute_dtype)
    output_image, count_map = None, None
    for slice_g in range(0, total_slices, sw_batch_size):
        slice_range = range(slice_g, min(slice_g + sw_batch_size, total_slices))
        unravel = [[slice(idx // num_win, idx // num_win + 1), slice(None)] + list(slices[idx % num_win])
                   for idx in slice_range]
        win_data = torch.cat([inputs[tuple(ws)] for ws in unravel])
        seg_prob = predictor(win_data)
        if isinstance(seg_prob, (tuple, list)):
            seg_prob = seg_prob[0]
        if output_image is None:
            out_shape = [batch_size, seg_prob.shape[1]] + list(image_size)
            output_image = torch.zeros(out_shape, dtype=compute_dtype, device=inputs.device)
            count_map = torch.zeros([1, 1] + list(image_size), dtype=compute_dtype, device=inputs.device)
            for ws in slices:
                count_map[(slice(None), slice(None)) + tuple(ws)] += importance_map
        w_t = importance_map.to(seg_prob.dtype)
        seg_prob = seg_prob * w_t
        for i, ws in enumerate(unravel):
            output_image[tuple(ws)] += seg_prob[i:i + 1].to(compute_dtype)
    output_image = output_image / count_map
    if any(pad_size):
        crop = [slice(None), slice(None)]
        for sp in range(num_spatial_dims):
            lo = pad_size[2 * (num_spatial_dims - 1 - sp)]
            crop.append(slice(lo, lo + image_size_[sp]))
        output_image = output_image[tuple(crop)]
    return output_image


# ------------------------------------------------ registration
def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _InertTransform:
    """Stand-in for a monai.transforms class that is constructed but never applied in the tests."""

    def __init__(self, *args, **kwargs):
        self.args = args
        self.__dict__.update(kwargs)

    def __call__(self, data):
        raise NotImplementedError(f"{type(self).__name__}: the MONAI shim only records this transform's arguments")


class _Compose(_InertTransform):
    def __init__(self, transforms=(), *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.transforms = list(transforms)

    def append(self, t):
        self.transforms.append(t)


class MapTransform:
    """monai.transforms.MapTransform as far as utils/gridmask.py uses it: stores the keys as a tuple [RECALLED]."""

    def __init__(self, keys, allow_missing_keys: bool = False):
        self.keys = (keys,) if isinstance(keys, str) else tuple(keys)
        self.allow_missing_keys = allow_missing_keys


def install():
    """Register the stand-in `monai`, `thop`, `timm`, `pyparsing` modules (idempotent)."""
    if "monai" in sys.modules and getattr(sys.modules["monai"], "__fcd_shim__", False):
        return
    _mod("monai", __fcd_shim__=True, __version__="1.5.1-shim")
    _mod("monai.utils", ensure_tuple_rep=ensure_tuple_rep, InterpolateMode=InterpolateMode, UpsampleMode=UpsampleMode,
         optional_import=lambda *a, **k: (None, False))
    _mod("monai.networks")
    _mod("monai.networks.layers")
    _mod("monai.networks.layers.factories", Conv=Conv, Norm=Norm, Act=Act, Dropout=Dropout, Pool=Pool, Pad=Pad)
    _mod("monai.networks.layers.utils", get_norm_layer=get_norm_layer, get_act_layer=get_act_layer)
    _mod("monai.networks.blocks", UpSample=UpSample)
    _mod("monai.networks.blocks.convolutions", Convolution=Convolution)
    _mod("monai.networks.blocks.dynunet_block", get_conv_layer=dyn_get_conv_layer, UnetOutBlock=UnetOutBlock,
         UnetResBlock=_Unavailable)
    _mod("monai.networks.blocks.mlp", MLPBlock=MLPBlock)
    _mod("monai.networks.blocks.upsample", UpSample=UpSample, SubpixelUpsample=SubpixelUpsample)
    _mod("monai.networks.blocks.segresnet_block", ResBlock=ResBlock, get_conv_layer=seg_get_conv_layer,
         get_upsample_layer=get_upsample_layer)
    _mod("monai.networks.utils", pixelshuffle=pixelshuffle)
    _mod("monai.networks.nets", SegResNet=SegResNet, SegResNetVAE=SegResNetVAE, UNETR=_Unavailable,
         SwinUNETR=_Unavailable, VNet=_Unavailable, UNet=_Unavailable)
    _mod("monai.losses", DiceLoss=DiceLoss, DiceCELoss=DiceCELoss, DiceFocalLoss=DiceFocalLoss, FocalLoss=FocalLoss,
         GeneralizedDiceLoss=GeneralizedDiceLoss, GeneralizedDiceFocalLoss=GeneralizedDiceFocalLoss)
    _mod("monai.inferers", sliding_window_inference=sliding_window_inference)
    # utils/gridmask.py:5 subclasses MapTransform for its dictionary wrapper; only `keys` is used (gridmask.py:126, 144)
    tr = _mod("monai.transforms", MapTransform=MapTransform, Compose=_Compose)
    # get_transforms.py:2-9 imports two dozen dictionary transforms by name only to build its pipelines; none of them is
    # on the path under test (the per-case file transforms are out of scope, the per-patch ones are restated by
    # oracle/sampling.py).  Any other name resolves to an inert record of its constructor arguments, which is all
    # FCDTrainTransform.set_prob touches (`self.coarse_dropout.prob`, get_transforms.py:116-120).
    tr.__getattr__ = lambda name: type(name, (_InertTransform,), {})
    if "thop" not in sys.modules:
        _mod("thop", profile=lambda *a, **k: (0, 0), clever_format=lambda x, *a, **k: x)
    if "pyparsing" not in sys.modules:
        _mod("pyparsing", Optional=object)
    if "timm" not in sys.modules:
        _mod("timm")
        _mod("timm.layers", trunc_normal_=nn.init.trunc_normal_)
        _mod("timm.models")
        _mod("timm.models.layers", trunc_normal_=nn.init.trunc_normal_)
