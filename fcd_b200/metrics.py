"""Voxel-level evaluation on the device: the reference's metrics.py:74-160 (`_compute_metrics`,
`calculate_voxel_level_metrics`, called from ModelTrainer.evaluate, train.py:220), seg_fcd_test.py:160-178 (per-subject Dice /
IoU) and utils/utils_common.py:37-60 (`evaluate_fp`) without moving the volumes to the host: one streaming counting
kernel per call (csrc/metrics.cu), the ratios on the [B, C, 4] count table.  The reference's optional ROC-AUC and HD95
branches (off at train.py:220 unless `include_hd95`) are not built.

MONAI's conventions are kept [RECALLED, MONAI 1.5.1]: include_background=False drops channel 0 only when there are several
channels; precision / sensitivity / F1 are ratios of the MEAN confusion table (channels, then batch); Dice is NaN for an
empty ground truth and NaNs are left out of the mean (0 when nothing remains)."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call


def confusion_counts(pred: torch.Tensor, label: torch.Tensor, thr_pred: float = 0.5, thr_label: float = 0.5):
    """[B, C, 4] int64 device tensor: tp, fp, tn, fn of (pred > thr_pred) vs (label > thr_label) per (batch, channel).
    pred: [B, C, ...] fp32 / uint8 / bool CUDA tensor, label: same shape (any float or integer type)."""
    if not pred.is_cuda or not label.is_cuda:
        raise RuntimeError("fcd_b200.metrics runs on CUDA tensors only (no CPU fallback)")
    if pred.shape != label.shape or pred.dim() < 3:
        raise ValueError(f"pred {tuple(pred.shape)} and label {tuple(label.shape)} must both be [B, C, spatial...]")
    B, C = pred.shape[:2]
    n = pred[0, 0].numel()
    pf = pu = None
    if pred.dtype in (torch.uint8, torch.bool):
        pu = pred.contiguous().view(torch.uint8)
    else:
        pf = pred.float().contiguous()
    lab = label.float().contiguous()
    counts = torch.empty((B, C, 4), dtype=torch.int64, device=pred.device)
    call("fcd_confusion_counts", pred_f=pf, pred_u8=pu, label=lab, thr_pred=float(thr_pred), thr_label=float(thr_label),
         n=n, items=B * C, counts=counts)
    return counts


def _drop_background(counts):
    return counts[:, 1:] if counts.shape[1] > 1 else counts


def _nanmean_channels_then_batch(f):
    """monai.metrics.utils.do_metric_reduction(f, 'mean') on a [B, C] table (device, no synchronisation)."""
    ok = ~torch.isnan(f)
    f = torch.where(ok, f, torch.zeros_like(f))
    nc = ok.sum(1)
    per_b = torch.where(nc > 0, f.sum(1) / nc.clamp(min=1), torch.zeros_like(f[:, 0]))
    nb = (nc > 0).sum()
    return torch.where(nb > 0, per_b.sum() / nb.clamp(min=1), torch.zeros_like(per_b.sum()))


def _ratio(a, b):
    return torch.where(b != 0, a / b, torch.full_like(a, float("nan")))


def metrics_from_counts(counts: torch.Tensor, as_tensors: bool = False) -> dict:
    """The dictionary of `_compute_metrics` (metrics.py:97-104) from a [B, C, 4] confusion table."""
    c = _drop_background(counts).double()
    tp, fp, fn = c[..., 0], c[..., 1], c[..., 3]
    dice = torch.where(tp + fn > 0, 2 * tp / (2 * tp + fp + fn).clamp(min=1), torch.full_like(tp, float("nan")))
    m = c.mean(1).mean(0)
    vals = torch.stack([_ratio(m[0], m[0] + m[1]), _ratio(m[0], m[0] + m[3]), _ratio(2 * m[0], 2 * m[0] + m[3] + m[1]),
                        _nanmean_channels_then_batch(dice)])
    keys = ("Prec", "Sens", "F1", "DC")
    if as_tensors:
        return dict(zip(keys, vals.unbind(0)))
    return dict(zip(keys, vals.tolist()))


def compute_metrics(y_pred: torch.Tensor, y_true: torch.Tensor, compute_roc_auc: bool = False,
                    compute_hd95: bool = False, as_tensors: bool = False) -> dict:
    """metrics.py:74-126 `_compute_metrics`: {'Prec', 'Sens', 'F1', 'DC'} of [B, C, ...] tensors binarised at 0.5.
    Python floats as the reference returns (ONE device->host copy of 4 doubles), or 0-d device tensors with
    as_tensors=True (no synchronisation)."""
    if compute_roc_auc or compute_hd95:
        raise NotImplementedError("ROC-AUC / HD95 (metrics.py:108-121) are not built: evaluation extras outside SURVEY 8")
    return metrics_from_counts(confusion_counts(y_pred, y_true), as_tensors)


class VoxelMetricAccumulator:
    """The global branch of calculate_voxel_level_metrics (metrics.py:156-160) without keeping the volumes: the reference
    concatenates every subject's prediction and label and counts per (subject, channel); the per-subject count tables
    are all that computation reads, so `update` keeps 4 integers per subject and channel (subjects may differ in size,
    which torch.cat in the reference cannot take)."""

    def __init__(self):
        self.tables = []

    def update(self, pred: torch.Tensor, label: torch.Tensor) -> None:
        lift = lambda t: t[None, None] if t.dim() == 3 else t
        self.tables.append(confusion_counts(lift(pred), lift(label)))

    def aggregate(self, as_tensors: bool = False) -> dict:
        if not self.tables:
            raise RuntimeError("VoxelMetricAccumulator.aggregate: no subject was added")
        return metrics_from_counts(torch.cat(self.tables), as_tensors)


def calculate_voxel_level_metrics(predictions, labels, compute_roc_auc: bool = False, compute_hd95: bool = False,
                                  average_across_subjects: bool = False) -> dict:
    """metrics.py:128-160: lists of per-subject [D,H,W] (or [1,C,D,H,W]) CUDA volumes, as ModelTrainer.evaluate collects
    them (train.py:214-220)."""
    lift = lambda t: t[None, None] if t.dim() == 3 else t
    if average_across_subjects:
        per = [compute_metrics(lift(p), lift(l), compute_roc_auc, compute_hd95, as_tensors=True)
               for p, l in zip(predictions, labels)]
        tab = torch.stack([torch.stack([m[k] for k in per[0]]) for m in per]).mean(0)
        return dict(zip(per[0].keys(), tab.tolist()))
    return compute_metrics(torch.cat([lift(p) for p in predictions]), torch.cat([lift(l) for l in labels]),
                           compute_roc_auc, compute_hd95)


def dice_iou(pred: torch.Tensor, label: torch.Tensor):
    """seg_fcd_test.py:160-178 for one subject: (dice, iou) floats of [B, C, ...] binary tensors; an empty ground truth
    scores 1 / 1 against an empty prediction and 0 / 0 otherwise."""
    # binary {0, 1} tensors (the reference passes AsDiscrete outputs): one counting pass serves both the edge case,
    # decided on ALL channels (labels.sum() / pred.sum()), and the metrics of the foreground channels
    allc = confusion_counts(pred, label)
    tot_t = (allc[..., 0] + allc[..., 3]).sum()
    tot_p = (allc[..., 0] + allc[..., 1]).sum()
    c = _drop_background(allc).double()
    tp, fp, fn = c[..., 0], c[..., 1], c[..., 3]
    nan = torch.full_like(tp, float("nan"))
    dice = _nanmean_channels_then_batch(torch.where(tp + fn > 0, 2 * tp / (2 * tp + fp + fn).clamp(min=1), nan))
    iou = _nanmean_channels_then_batch(torch.where(tp + fn > 0, tp / (tp + fp + fn).clamp(min=1), nan))
    edge = torch.where(tot_p == 0, torch.ones_like(dice), torch.zeros_like(dice))
    out = torch.stack([torch.where(tot_t == 0, edge, dice), torch.where(tot_t == 0, edge, iou)]).tolist()
    return out[0], out[1]


_FP_WS = {}


def evaluate_fp(cc: torch.Tensor, label: torch.Tensor, max_id: int | None = None) -> torch.Tensor:
    """utils/utils_common.py:37-60: the number of connected components of `cc` (ids > 0, e.g. the second output of
    fcd_b200.post_process_segment) that share no voxel with a non-zero `label`.  0-d int64 device tensor, no host
    synchronisation.  max_id: an upper bound of the ids (default: the voxel count)."""
    if not cc.is_cuda or not label.is_cuda:
        raise RuntimeError("fcd_b200.metrics runs on CUDA tensors only (no CPU fallback)")
    if cc.numel() != label.numel():
        raise ValueError("cc and label must hold one value per voxel")
    V = cc.numel()
    max_id = int(max_id or V)
    key = (max_id, cc.device.index)
    ws = _FP_WS.get(key)
    if ws is None:
        if len(_FP_WS) >= 4:
            _FP_WS.clear()
        ws = _FP_WS[key] = torch.empty(_lib.lib().fcd_component_overlap_ws_bytes(max_id), dtype=torch.uint8,
                                       device=cc.device)
    out = torch.empty((3,), dtype=torch.int64, device=cc.device)
    call("fcd_component_overlap", cc=cc.float().contiguous(), label=label.float().contiguous(), V=V, max_id=max_id,
         ws=ws, ws_bytes=ws.numel(), out=out)
    # ids outside [0, max_id] cannot be counted: poison the result instead of returning a wrong number silently
    return torch.where(out[2] == 0, out[0] - out[1], torch.full_like(out[0], -1))
