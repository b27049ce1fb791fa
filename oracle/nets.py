"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the network forward passes (plain PyTorch fp32, functional).

A restatement, driven by a reference-format ``state_dict``, of the forward arithmetic of the networks
``get_model`` builds (reference get_model.py:9-249).  Each function cites the reference file:line it
follows (paths relative to the reference root).  MONAI 1.5.1 pieces are restated from its published
behaviour because MONAI is absent from this image; for those pieces parity is
UNPINNED against a real MONAI install.  tests/test_oracle_vs_reference.py checks every function here
against the reference's own source files (run through oracle/monai_shim.py) whenever /root/reference
is present, and tests/golden/ holds outputs of those reference files for the GPU box.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  All Dropout layers are evaluated with p = 0 (parity runs; RNG streams cannot match).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

LRELU = 0.01
EPS = 1e-5


# ----------------------------------------------------------------------------------------------- blocks
def _norm(sd, pre, x, kind, training, bn_out=None):
    """get_norm_layer targets: InstanceNorm3d(affine=False) / BatchNorm3d (conv_blocks.py:418-419,437)."""
    if kind == "instance":
        return F.instance_norm(x, eps=EPS)
    if kind == "batch":
        rm, rv = sd[pre + ".running_mean"].clone(), sd[pre + ".running_var"].clone()
        y = F.batch_norm(x, rm, rv, sd[pre + ".weight"], sd[pre + ".bias"], training=training, momentum=0.1, eps=EPS)
        if training and bn_out is not None:
            bn_out[pre + ".running_mean"] = rm
            bn_out[pre + ".running_var"] = rv
            bn_out[pre + ".num_batches_tracked"] = sd[pre + ".num_batches_tracked"] + 1
        return y
    raise ValueError(kind)


def unet_res_block(sd, pre, x, norm="instance", training=True, bn_out=None):
    """UnetResBlock.forward (networks/ms_dsa_net/conv_blocks.py:439-452)."""
    out = F.conv3d(x, sd[pre + ".conv1.conv.weight"], padding=1)
    out = F.leaky_relu(_norm(sd, pre + ".norm1", out, norm, training, bn_out), LRELU)
    out = F.conv3d(out, sd[pre + ".conv2.conv.weight"], padding=1)
    out = _norm(sd, pre + ".norm2", out, norm, training, bn_out)
    res = x
    if (pre + ".conv3.conv.weight") in sd:
        res = F.conv3d(x, sd[pre + ".conv3.conv.weight"])
        res = _norm(sd, pre + ".norm3", res, norm, training, bn_out)
    return F.leaky_relu(out + res, LRELU)


def unetr_up_block(sd, pre, inp, skip):
    """UnetrUpBlock.forward (conv_blocks.py:681-689): ConvTranspose3d k2 s2 -> cat -> UnetResBlock."""
    out = F.conv_transpose3d(inp, sd[pre + ".transp_conv.conv.weight"], stride=2)
    return unet_res_block(sd, pre + ".conv_block", torch.cat((out, skip), 1))


def subpixel_upsample(sd, pre, x):
    """MONAI SubpixelUpsample: conv3 (bias) -> pixelshuffle x2 -> pad(1,0)x3 -> AvgPool3d(2,1)."""
    x = F.conv3d(x, sd[pre + ".conv_block.weight"], sd[pre + ".conv_block.bias"], padding=1)
    b, c8, d, h, w = x.shape
    c = c8 // 8
    x = x.reshape(b, c, 2, 2, 2, d, h, w).permute(0, 1, 5, 2, 6, 3, 7, 4).reshape(b, c, 2 * d, 2 * h, 2 * w)
    x = F.pad(x, (1, 0, 1, 0, 1, 0))
    return F.avg_pool3d(x, 2, 1)


def upsample(sd, pre, x):
    """MONAI UpSample, mode read off the children present in the state dict: pixelshuffle | deconv
    (ConvTranspose3d k2 s2, bias) | nontrainable (optional 1x1 preconv + trilinear x2, align_corners=False)."""
    if (pre + ".pixelshuffle.conv_block.weight") in sd:
        return subpixel_upsample(sd, pre + ".pixelshuffle", x)
    if (pre + ".deconv.weight") in sd:
        return F.conv_transpose3d(x, sd[pre + ".deconv.weight"], sd.get(pre + ".deconv.bias"), stride=2)
    if (pre + ".preconv.weight") in sd:
        x = F.conv3d(x, sd[pre + ".preconv.weight"], sd.get(pre + ".preconv.bias"))
    return F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=False)


def general_up_block(sd, pre, inp, skip):
    """GeneralUnetrUpBlock.forward (conv_blocks.py:766-775)."""
    out = upsample(sd, pre + ".upsample", inp)
    return unet_res_block(sd, pre + ".conv_block", torch.cat((out, skip), 1))


def dsa(sd, pre, x, heads=4):
    """DSA.forward (conv_blocks.py:316-355), dropout p=0.  x: [B,N,C].  The sa_type is read off the projection count of
    qkvv.weight: 4C rows = 'parallel' (328-355); 3C rows = 'spatial' (forward_spatial, 236-258), 'channel'
    (forward_channel, 260-279) or 'serial' (forward_serial, 281-314: the spatial output, NOT scrambled, is the value of
    the channel attention), told apart by the key `<pre>.__sa_type__` the tests put into the state dict ('spatial' if
    absent)."""
    B, N, C = x.shape
    c = C // heads
    W = sd[pre + ".qkvv.weight"]
    if W.shape[0] == 3 * C:
        qkv = F.linear(x, W).reshape(B, N, 3, heads, c).permute(2, 0, 3, 1, 4)
        q, k, v = (t.transpose(-2, -1) for t in (qkv[0], qkv[1], qkv[2]))                  # [B,h,c,N]
        if sd.get(pre + ".__sa_type__", "spatial") == "channel":
            q, k = F.normalize(q, dim=-1), F.normalize(k, dim=-1)
            attn = ((q @ k.transpose(-2, -1)) * sd[pre + ".temperature"]).softmax(dim=-1)
            return (attn @ v).permute(0, 3, 1, 2).reshape(B, N, C)
        EF = sd[pre + ".EF"]
        k_proj = torch.einsum("bhdn,nk->bhdk", k, EF)
        v_proj = torch.einsum("bhdn,nk->bhdk", v, EF)
        q, k = F.normalize(q, dim=-1), F.normalize(k, dim=-1)
        attn = ((q.permute(0, 1, 3, 2) @ k_proj) * sd[pre + ".temperature2"]).softmax(dim=-1)
        if sd.get(pre + ".__sa_type__", "spatial") == "serial":
            x_sa = attn @ v_proj.transpose(-2, -1)                                          # [B,h,N,c]
            attn_ca = ((q @ k.transpose(-2, -1)) * sd[pre + ".temperature"]).softmax(dim=-1)
            return (attn_ca @ x_sa.transpose(-2, -1)).permute(0, 3, 1, 2).reshape(B, N, C)
        return (attn @ v_proj.transpose(-2, -1)).permute(0, 3, 1, 2).reshape(B, N, C)
    qkvv = F.linear(x, W).reshape(B, N, 4, heads, c).permute(2, 0, 3, 1, 4)
    q, k, v_ca, v_sa = (t.transpose(-2, -1) for t in (qkvv[0], qkvv[1], qkvv[2], qkvv[3]))  # [B,h,c,N]
    EF = sd[pre + ".EF"]
    k_proj = torch.einsum("bhdn,nk->bhdk", k, EF)
    v_proj = torch.einsum("bhdn,nk->bhdk", v_sa, EF)
    q = F.normalize(q, dim=-1)
    k = F.normalize(k, dim=-1)
    attn_ca = ((q @ k.transpose(-2, -1)) * sd[pre + ".temperature"]).softmax(dim=-1)
    x_ca = (attn_ca @ v_ca).permute(0, 3, 1, 2).reshape(B, N, C)
    attn_sa = ((q.permute(0, 1, 3, 2) @ k_proj) * sd[pre + ".temperature2"]).softmax(dim=-1)
    # line 353: [B,h,N,c] -> permute(0,3,1,2) = [B,c,h,N] -> reshape(B,N,C): a memory scramble, kept as is
    x_sa = (attn_sa @ v_proj.transpose(-2, -1)).permute(0, 3, 1, 2).reshape(B, N, C)
    return x_ca + x_sa


def with_sa_type(sd, sa_type):
    """Copy of the state dict carrying the `<prefix>.dsa.__sa_type__` markers `dsa` reads (3-projection types only)."""
    out = dict(sd)
    for k in sd:
        if k.endswith(".dsa.qkvv.weight"):
            out[k[:-len("qkvv.weight")] + "__sa_type__"] = sa_type
    return out


def transformer_block(sd, pre, x, training=True, bn_out=None):
    """TransformerBlock.forward (conv_blocks.py:69-90), Dropout3d p=0."""
    B, C, H, W, D = x.shape
    t = x.reshape(B, C, H * W * D).permute(0, 2, 1)
    if (pre + ".pos_embed") in sd:
        t = t + sd[pre + ".pos_embed"]
    ln = F.layer_norm(t, (C,), sd[pre + ".norm.weight"], sd[pre + ".norm.bias"], 1e-5)
    t = t + sd[pre + ".gamma"] * dsa(sd, pre + ".dsa", ln)
    x = t.reshape(B, H, W, D, C).permute(0, 4, 1, 2, 3)
    y = unet_res_block(sd, pre + ".conv51", x, norm="batch", training=training, bn_out=bn_out)
    y = F.conv3d(y, sd[pre + ".conv8.1.weight"], sd[pre + ".conv8.1.bias"])
    return x + y


def _patch_embed(sd, pre, x):
    """patch_embedding{3..6}: 1x1 conv (no bias) + GroupNorm(C/2 groups, affine) (ms_dsa_net.py:215-218)."""
    x = F.conv3d(x, sd[pre + ".0.conv.weight"])
    w = sd[pre + ".1.weight"]
    return F.group_norm(x, w.numel() // 2, w, sd[pre + ".1.bias"], EPS)


# ---------------------------------------------------------------------------------------------- networks
def base_unet(sd, x, depth=6):
    """BaseUNet.forward (networks/ms_dsa_net/ms_dsa_net.py:84-101)."""
    feats, out = [], x
    for i in range(depth):
        out = unet_res_block(sd, f"encoders.{i}.layer", out)
        feats.append(out)
        if i != depth - 1:
            out = F.max_pool3d(out, 2, 2)
    for i in range(depth - 1):
        out = unetr_up_block(sd, f"decoders.{i}", out, feats[-(i + 2)])
    return F.conv3d(out, sd["final_conv.weight"], sd["final_conv.bias"])


def ms_dsa_net(sd, x, training=True, bn_out=None, pixelshuffle=False):
    """MS_DSA_NET.forward (ms_dsa_net.py:375-407) / MS_DSA_NET_PS.forward (same flow, pixelshuffle decoders)."""
    xs = []
    out = x
    for i in range(1, 7):
        out = unet_res_block(sd, f"encoder{i}.layer", out if i == 1 else F.max_pool3d(out, 2, 2))
        xs.append(out)
    ts = {}
    for lvl in (3, 4, 5, 6):
        t = _patch_embed(sd, f"patch_embedding{lvl}", xs[lvl - 1])
        for j in range(3):
            t = transformer_block(sd, f"trans{lvl}.{j}", t, training, bn_out)
        ts[lvl] = t
    up = general_up_block if pixelshuffle else unetr_up_block
    y = up(sd, "decoder5", ts[6], ts[5])
    y = up(sd, "decoder4", y, ts[4])
    y = up(sd, "decoder3", y, ts[3])
    y = up(sd, "decoder2", y, xs[1])
    y = up(sd, "decoder1", y, xs[0])
    return F.conv3d(y, sd["out.conv.conv.weight"], sd["out.conv.conv.bias"])


def _seg_resblock(sd, pre, x):
    """MONAI ResBlock: IN-ReLU-conv3-IN-ReLU-conv3 + identity."""
    y = F.conv3d(F.relu(F.instance_norm(x, eps=EPS)), sd[pre + ".conv1.conv.weight"], padding=1)
    y = F.conv3d(F.relu(F.instance_norm(y, eps=EPS)), sd[pre + ".conv2.conv.weight"], padding=1)
    return y + x


def _seg_up(sd, pre, x):
    """up_samples[i] = 1x1 conv C->C/2 + UpSample(pixelshuffle) (segresnet_dsa.py:130-141)."""
    x = F.conv3d(x, sd[pre + ".0.conv.weight"])
    return upsample(sd, pre + ".1", x)


def _seg_final(sd, pre, x):
    """_make_final_conv: IN -> ReLU -> 1x1 conv with bias (segresnet_dsa.py:188-193)."""
    return F.conv3d(F.relu(F.instance_norm(x, eps=EPS)), sd[pre + ".2.conv.weight"], sd[pre + ".2.conv.bias"])


def _count(sd, fmt):
    n = 0
    while any(k.startswith(fmt.format(n)) for k in sd):
        n += 1
    return n


def segresnet(sd, x, training=True, bn_out=None, vae_noise=None, vae=False):
    """SegResNet / SegResNetVAE (MONAI) and SegResNet_DSA / SegResNetVAE_DSA
    (networks/segresnet_dsa/segresnet_dsa.py:195-230, 322-373).  Dropout3d p=0.

    Returns logits, or (logits, vae_loss | None) when ``vae``.  ``vae_noise`` replaces randn_like (line 332).
    """
    net_input = x
    x = F.conv3d(x, sd["convInit.conv.weight"], padding=1)
    n_down = _count(sd, "down_layers.{}.")
    n_trans = _count(sd, "transformer_layers.{}.")
    dsa_start_level = (n_down - n_trans) if n_trans else None   # get_model.py:192: len(blocks_down) - 2
    down_x = []
    feature = None
    for i in range(n_down):
        if i > 0:
            x = F.conv3d(x, sd[f"down_layers.{i}.0.conv.weight"], stride=2, padding=1)
        j = 1
        while f"down_layers.{i}.{j}.conv1.conv.weight" in sd:
            x = _seg_resblock(sd, f"down_layers.{i}.{j}", x)
            j += 1
        feature = x
        if dsa_start_level is not None and i >= dsa_start_level:
            li = i - dsa_start_level
            for blk in range(_count(sd, f"transformer_layers.{li}." + "{}.")):
                feature = transformer_block(sd, f"transformer_layers.{li}.{blk}", feature, training, bn_out)
        down_x.append(feature)
    x = feature
    down_x.reverse()
    vae_input = x
    n_up = _count(sd, "up_samples.{}.")

    def run_up(x, skips):
        for i in range(n_up):
            x = _seg_up(sd, f"up_samples.{i}", x)
            if skips is not None:
                x = x + skips[i + 1]
            j = 0
            while f"up_layers.{i}.{j}.conv1.conv.weight" in sd:
                x = _seg_resblock(sd, f"up_layers.{i}.{j}", x)
                j += 1
        return x

    logits = _seg_final(sd, "conv_final", run_up(x, down_x))
    if not vae:
        return logits
    if not training:
        return logits, None
    v = F.relu(F.instance_norm(vae_input, eps=EPS))
    v = F.conv3d(v, sd["vae_down.2.conv.weight"], sd["vae_down.2.conv.bias"], stride=2, padding=1)
    v = F.relu(F.instance_norm(v, eps=EPS))
    fc_shape = v.shape[1:]
    v = v.reshape(-1, sd["vae_fc1.weight"].shape[1])
    z_mean = F.linear(v, sd["vae_fc1.weight"], sd["vae_fc1.bias"])
    noise = vae_noise if vae_noise is not None else torch.randn_like(z_mean)
    reg = torch.mean(z_mean ** 2)
    v = z_mean + 0.3 * noise
    v = F.relu(F.linear(v, sd["vae_fc3.weight"], sd["vae_fc3.bias"]))
    v = v.reshape(-1, *fc_shape)
    v = F.conv3d(v, sd["vae_fc_up_sample.0.conv.weight"])
    v = upsample(sd, "vae_fc_up_sample.1", v)
    v = F.relu(F.instance_norm(v, eps=EPS))
    v = _seg_final(sd, "vae_conv_final", run_up(v, None))
    return logits, reg + F.mse_loss(net_input, v)


def forward(model_type, sd, x, training=True, bn_out=None, vae_noise=None):
    """Dispatch mirroring get_model.py:9-249 for the in-scope model types."""
    mt = model_type.lower()
    if mt == "baseunet":
        return base_unet(sd, x)
    if mt == "ms_dsa_net":
        return ms_dsa_net(sd, x, training, bn_out)
    if mt == "ms_dsa_net_ps":
        return ms_dsa_net(sd, x, training, bn_out, pixelshuffle=True)
    if mt == "segresnet":
        return segresnet(sd, x, training, bn_out)
    if mt == "segresnetvae":
        return segresnet(sd, x, training, bn_out, vae_noise, vae=True)
    if mt == "segresnet_dsa":
        return segresnet(sd, x, training, bn_out)
    if mt == "segresnetvae_dsa":
        return segresnet(sd, x, training, bn_out, vae_noise, vae=True)
    raise ValueError(model_type)
