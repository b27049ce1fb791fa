"""CombinedLoss with the reference's constructor / forward signature (get_loss.py:10-39), computed by ONE fused
CUDA reduction (fcd_b200/csrc/loss.cu) instead of MONAI's DiceLoss / DiceCELoss / DiceFocalLoss / GeneralizedDice(Focal)Loss + ATen slicing."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


def loss_config(params: dict) -> dict:
    """get_loss_function_from_params (get_loss.py:42-97): every loss type the reference builds."""
    kind = params.get("loss", "DiceLoss")
    if kind not in ops.LOSS_KIND:
        raise NotImplementedError(f"loss {kind!r}: the reference builds only {sorted(ops.LOSS_KIND)} (get_loss.py:54-95)")
    wtype = str(params.get("gdice_wtype", "square"))
    if wtype not in ops.GDICE_WTYPE:
        raise ValueError(f"gdice_wtype {wtype!r}: expected one of {sorted(ops.GDICE_WTYPE)} (config.py:45)")
    if params.get("sigmoid", False) or not params.get("softmax", True) or params.get("chans_out", 2) != 2:
        raise NotImplementedError("fused loss implements the reference default: softmax=True, sigmoid=False, chans_out=2")
    lam2 = params.get("lambda_ce", 1.0) if kind == "DiceCELoss" else params.get("lambda_focal", 1.0)
    return dict(kind=ops.LOSS_KIND[kind], lambda_dice=float(params.get("lambda_dice", 1.0)), lambda_2=float(lam2),
                w_bg=float(params.get("ce_background_weight", 0.5)), w_fg=float(params.get("ce_fcd_weight", 0.5)),
                gamma=float(params.get("gamma_focal", 2.0)), squared=int(bool(params.get("square_pred", False))),
                jaccard=int(bool(params.get("jaccard", False))), w_type=ops.GDICE_WTYPE[wtype], smooth_nr=1e-5, smooth_dr=1e-5,
                tv_w=float(params.get("tv_loss_weight", 0.0)),
                tv_norm=2 if params.get("tv_loss_norm", "l1") == "l2" else 1,
                tv_exclude=int(bool(params.get("tvloss_exclude_borders", False))))


class CombinedLoss(nn.Module):
    def __init__(self, params: dict, device) -> None:
        super().__init__()
        self.params = params
        self.device = device
        self.tv_loss_weight = params.get("tv_loss_weight", 0.0)
        self.boundaryloss_weight = params.get("boundaryloss_weight", 0.0)
        self.caloss_weight = params.get("caloss_weight", 0.0)
        if self.boundaryloss_weight > 0 or self.caloss_weight > 0:
            raise NotImplementedError("boundary / cortical terms are dead with the reference config (SURVEY 2.6)")
        self.cfg = loss_config(params)

    def forward(self, pred: torch.Tensor, target: torch.Tensor, thickness_map=None) -> torch.Tensor:
        if not pred.is_cuda:
            raise RuntimeError("fcd_b200 CombinedLoss runs on CUDA only; there is no CPU fallback")
        return ops.fused_loss(pred, target, self.cfg)


def get_loss_function_from_params(params, device):
    return CombinedLoss(params, device)
