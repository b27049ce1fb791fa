"""TEST INFRASTRUCTURE ONLY -- CPU oracle for CombinedLoss (reference get_loss.py:10-165), plain PyTorch fp32.

Restates MONAI 1.5.1 DiceLoss / DiceCELoss / DiceFocalLoss / FocalLoss semantics for the exact kwargs
get_loss_function_from_params passes (get_loss.py:46-78: include_background=False, to_onehot_y=True,
softmax=True, batch=True, smooth 1e-5) -- SURVEY.md Appendix A6, parity UNPINNED against real MONAI --
and the reference's own compute_total_variation_loss / dilate_mask (get_loss.py:100-165).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _fg_prob_and_onehot(pred, target):
    p1 = torch.softmax(pred.float(), dim=1)[:, 1:2]
    y = (target.float() == 1).float()          # one_hot(target, 2)[:, 1:]
    return p1, y


def dice(pred, target, smooth_nr=1e-5, smooth_dr=1e-5, squared_pred=False, jaccard=False):
    """MONAI DiceLoss, foreground channel only, reduced over batch+space (get_loss.py:46-57)."""
    p1, y = _fg_prob_and_onehot(pred, target)
    inter = (p1 * y).sum()
    if squared_pred:
        den = (y ** 2).sum() + (p1 ** 2).sum()
    else:
        den = y.sum() + p1.sum()
    if jaccard:
        den = 2.0 * (den - inter)
    return 1.0 - (2.0 * inter + smooth_nr) / (den + smooth_dr)


def cross_entropy(pred, target, w_bg=0.5, w_fg=0.5):
    """nn.CrossEntropyLoss(weight=[w_bg,w_fg], reduction='mean') on logits (get_loss.py:59-69)."""
    w = torch.tensor([w_bg, w_fg], dtype=torch.float32, device=pred.device)
    return F.cross_entropy(pred.float(), target.squeeze(1).long(), weight=w)


def focal(pred, target, gamma=2.0):
    """MONAI FocalLoss(include_background=False, use_softmax=False): sigmoid focal on the foreground logit."""
    x = pred.float()[:, 1:2]
    t = (target.float() == 1).float()
    bce = x - x * t - F.logsigmoid(x)
    loss = torch.exp(gamma * F.logsigmoid(-x * (t * 2 - 1))) * bce
    return loss.mean(dim=(2, 3, 4)).mean()


def _dilate(mask, iterations=2):
    """dilate_mask (get_loss.py:100-113): 3^3 all-ones conv > 0, repeated."""
    k = torch.ones((1, 1, 3, 3, 3), device=mask.device)
    d = mask
    for _ in range(iterations):
        d = (F.conv3d(d.float(), k, padding=1) > 0).float()
    return d


def total_variation(pred, gt, norm=1, exclude_borders=False):
    """compute_total_variation_loss (get_loss.py:116-165) with softmax=True, sigmoid=False."""
    p = torch.softmax(pred.float(), dim=1)[:, 1:2]
    if exclude_borders:
        border = ((_dilate(gt) - (1 - _dilate(1 - gt))) > 0).float()
        p = p * (1 - border)
    dz = p[:, :, 1:] - p[:, :, :-1]
    dy = p[:, :, :, 1:] - p[:, :, :, :-1]
    dx = p[:, :, :, :, 1:] - p[:, :, :, :, :-1]
    if norm == 1:
        return dz.abs().mean() + dy.abs().mean() + dx.abs().mean()
    eps = 1e-10
    return sum(torch.sqrt((d ** 2).mean() + eps) for d in (dz, dy, dx))


def combined_loss(params, pred, target):
    """CombinedLoss.forward (get_loss.py:24-39) for loss in {DiceLoss, DiceCELoss, DiceFocalLoss} + TV."""
    kind = params.get("loss", "DiceLoss")
    d = dice(pred, target, squared_pred=params.get("square_pred", False), jaccard=params.get("jaccard", False))
    if kind == "DiceLoss":
        total = d
    elif kind == "DiceCELoss":
        total = params.get("lambda_dice", 1.0) * d + params.get("lambda_ce", 1.0) * cross_entropy(
            pred, target, params.get("ce_background_weight", 0.5), params.get("ce_fcd_weight", 0.5))
    elif kind == "DiceFocalLoss":
        total = params.get("lambda_dice", 1.0) * d + params.get("lambda_focal", 1.0) * focal(
            pred, target, params.get("gamma_focal", 2.0))
    else:
        raise ValueError(kind)
    tvw = params.get("tv_loss_weight", 0.0)
    if tvw > 0:
        total = total + tvw * total_variation(pred, target, norm=2 if params.get("tv_loss_norm") == "l2" else 1,
                                              exclude_borders=params.get("tvloss_exclude_borders", False))
    return total
