"""TEST INFRASTRUCTURE ONLY -- CPU oracle for CombinedLoss (reference get_loss.py:10-165), plain PyTorch fp32.

Restates MONAI 1.5.1 DiceLoss / DiceCELoss / DiceFocalLoss / FocalLoss / GeneralizedDice(Focal)Loss semantics for the exact kwargs
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


def generalized_dice(pred, target, w_type="square", smooth_nr=1e-5, smooth_dr=1e-5):
    """MONAI GeneralizedDiceLoss(include_background=True, batch=True) as get_loss.py:79-84 builds it: both softmax
    channels against the one-hot label, class weights from the label volumes, infinite weights -> the largest finite."""
    p = torch.softmax(pred.float(), dim=1)
    y = torch.cat([(target.float() != 1).float(), (target.float() == 1).float()], dim=1)
    ax = (0, 2, 3, 4)
    inter, g, s = (p * y).sum(ax), y.sum(ax), p.sum(ax)
    w = {"square": lambda v: 1.0 / (v * v), "simple": lambda v: 1.0 / v, "uniform": torch.ones_like}[w_type](g)
    inf = torch.isinf(w)
    w = torch.where(inf, torch.zeros_like(w), w)
    w = w + inf * w.max()
    return 1.0 - (2.0 * (inter * w).sum() + smooth_nr) / (((g + s) * w).sum() + smooth_dr)


def cross_entropy(pred, target, w_bg=0.5, w_fg=0.5):
    """nn.CrossEntropyLoss(weight=[w_bg,w_fg], reduction='mean') on logits (get_loss.py:59-69)."""
    w = torch.tensor([w_bg, w_fg], dtype=torch.float32, device=pred.device)
    return F.cross_entropy(pred.float(), target.squeeze(1).long(), weight=w)


def focal(pred, target, gamma=2.0, include_background=False):
    """MONAI FocalLoss(use_softmax=False): sigmoid focal on the foreground logit (include_background=False,
    get_loss.py:70-78) or on both logits against the one-hot label (include_background=True, get_loss.py:85-93)."""
    if include_background:
        x = pred.float()
        t = torch.cat([(target.float() != 1).float(), (target.float() == 1).float()], dim=1)
    else:
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
    """CombinedLoss.forward (get_loss.py:24-39) for every loss get_loss_function_from_params builds + TV."""
    kind = params.get("loss", "DiceLoss")
    if kind.startswith("Generalized"):
        d = generalized_dice(pred, target, w_type=params.get("gdice_wtype", "square"))
    else:
        d = dice(pred, target, squared_pred=params.get("square_pred", False), jaccard=params.get("jaccard", False))
    if kind in ("DiceLoss", "GeneralizedDiceLoss"):
        total = d
    elif kind == "DiceCELoss":
        total = params.get("lambda_dice", 1.0) * d + params.get("lambda_ce", 1.0) * cross_entropy(
            pred, target, params.get("ce_background_weight", 0.5), params.get("ce_fcd_weight", 0.5))
    elif kind == "DiceFocalLoss":
        total = params.get("lambda_dice", 1.0) * d + params.get("lambda_focal", 1.0) * focal(
            pred, target, params.get("gamma_focal", 2.0))
    elif kind == "GeneralizedDiceFocalLoss":
        total = params.get("lambda_dice", 1.0) * d + params.get("lambda_focal", 1.0) * focal(
            pred, target, params.get("gamma_focal", 2.0), include_background=True)
    else:
        raise ValueError(kind)
    tvw = params.get("tv_loss_weight", 0.0)
    if tvw > 0:
        total = total + tvw * total_variation(pred, target, norm=2 if params.get("tv_loss_norm") == "l2" else 1,
                                              exclude_borders=params.get("tvloss_exclude_borders", False))
    return total
