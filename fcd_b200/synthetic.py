"""Synthetic training / inference inputs of the reference's shapes (train.py:367-370: fp32 NCDHW image with
chans_in modalities, fp32 {0,1} lesion mask with ~1 % foreground).  Used by bench.py and smoke(); seeded torch RNG."""
from __future__ import annotations

import torch


def make_batch(batch, chans, size, seed=0, device="cpu", pin=False):
    if isinstance(size, int):
        size = (size,) * 3
    g = torch.Generator(device="cpu").manual_seed(1234 + seed)
    x = torch.randn((batch, chans) + tuple(size), generator=g)
    D, H, W = size
    zz = torch.arange(D).view(D, 1, 1)
    yy = torch.arange(H).view(1, H, 1)
    xx = torch.arange(W).view(1, 1, W)
    y = torch.zeros((batch, 1, D, H, W))
    for b in range(batch):
        for _ in range(3):
            c = torch.rand(3, generator=g) * 0.7 + 0.15
            r = torch.rand(3, generator=g) * 0.08 + 0.05
            m = ((zz - c[0] * D) / (r[0] * D)) ** 2 + ((yy - c[1] * H) / (r[1] * H)) ** 2 + \
                ((xx - c[2] * W) / (r[2] * W)) ** 2 <= 1.0
            y[b, 0][m] = 1.0
    if pin:
        x, y = x.pin_memory(), y.pin_memory()
    if device != "cpu":
        x, y = x.to(device), y.to(device)
    return x, y


def initialize_weights(module):
    """train_utils.py:44-60 -- what ModelTrainer applies after get_model (train.py:59)."""
    import torch.nn as nn
    if isinstance(module, (nn.Conv2d, nn.Conv3d)):
        nn.init.kaiming_normal_(module.weight, mode="fan_out", nonlinearity="relu")
        if module.bias is not None:
            nn.init.constant_(module.bias, 0)
    elif isinstance(module, nn.Linear):
        nn.init.xavier_uniform_(module.weight)
        if module.bias is not None:
            nn.init.constant_(module.bias, 0)
    elif isinstance(module, (nn.BatchNorm2d, nn.BatchNorm3d, nn.LayerNorm)):
        nn.init.constant_(module.weight, 1)
        nn.init.constant_(module.bias, 0)
