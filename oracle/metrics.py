"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's voxel-level evaluation (metrics.py:74-160) and of
utils/utils_common.py:37-60 `evaluate_fp`.

`evaluate_fp` is pinned against the reference's own function imported live (tests/test_oracle_vs_reference.py).  The
MONAI metric classes the reference calls are restated [RECALLED, MONAI 1.5.1; parity UNPINNED against a real MONAI, which
is not installable here]:
  get_confusion_matrix        : per (batch, channel) tp, fp, tn, fn of binarised tensors, summed over space
  ConfusionMatrixMetric("mean", compute_sample=False).aggregate(): the [B, C, 4] table is reduced to [4] first (mean over
                                channels, then over the batch), THEN precision = tp / (tp + fp), sensitivity = tp / (tp + fn),
                                specificity = tn / (tn + fp), f1 = 2 tp / (2 tp + fn + fp), NaN where the denominator is 0
  DiceMetric(include_background=False, reduction="mean", ignore_empty=True): per (batch, channel) 2 |p & t| / (|p| + |t|),
                                NaN when |t| = 0; mean over channels then batch ignoring NaN, 0 when nothing is left
  MeanIoU (seg_fcd_test.py:73-178): |p & t| / (|p| + |t| - |p & t|), NaN when |t| = 0, same reduction
  include_background=False drops channel 0 only when there is more than one channel (monai.metrics.utils.ignore_background).
"""
from __future__ import annotations

import numpy as np


def confusion_counts(pred, label, thr_pred=0.5, thr_label=0.5):
    """[B, C, 4] int64: tp, fp, tn, fn of (pred > thr_pred) against (label > thr_label), summed over the spatial axes."""
    p = np.asarray(pred) > thr_pred
    t = np.asarray(label) > thr_label
    ax = tuple(range(2, p.ndim))
    return np.stack([(p & t).sum(ax), (p & ~t).sum(ax), (~p & ~t).sum(ax), (~p & t).sum(ax)], axis=-1).astype(np.int64)


def _drop_background(counts):
    return counts[:, 1:] if counts.shape[1] > 1 else counts


def _nanmean_channels_then_batch(f):
    """monai.metrics.utils.do_metric_reduction(f, "mean") on a [B, C] table."""
    f = np.asarray(f, np.float64).copy()
    ok = ~np.isnan(f)
    f[~ok] = 0.0
    nc = ok.sum(1)
    per_b = np.where(nc > 0, f.sum(1) / np.maximum(nc, 1), 0.0)
    nb = (nc > 0).sum()
    return float(per_b.sum() / nb) if nb > 0 else 0.0


def metrics_from_counts(counts):
    """The dictionary of metrics.py:74-104 from the [B, C, 4] count table ('Spec' is computed by the reference and then
    left out of its dict)."""
    c = _drop_background(np.asarray(counts)).astype(np.float64)
    tp, fp, tn, fn = (c[..., i] for i in range(4))
    denom = 2 * tp + fp + fn                                   # |p| + |t|
    with np.errstate(divide="ignore", invalid="ignore"):
        dice = np.where(tp + fn > 0, 2 * tp / denom, np.nan)
    mtp, mfp, mtn, mfn = c.mean(1).mean(0)                     # no NaN in a count table: plain means
    ratio = lambda a, b: float(a / b) if b != 0 else float("nan")
    return {"Prec": ratio(mtp, mtp + mfp), "Sens": ratio(mtp, mtp + mfn), "Spec": ratio(mtn, mtn + mfp),
            "F1": ratio(2 * mtp, 2 * mtp + mfn + mfp), "DC": _nanmean_channels_then_batch(dice)}


def compute_metrics(y_pred, y_true):
    """metrics.py:74-126 `_compute_metrics` without the optional ROC-AUC / HD95 branches: {'Prec','Sens','F1','DC'}
    (+ 'Spec')."""
    return metrics_from_counts(confusion_counts(y_pred, y_true))


def calculate_voxel_level_metrics(predictions, labels, average_across_subjects=False):
    """metrics.py:128-160: lists of per-subject [D,H,W] (or [1,1,D,H,W]) volumes."""
    lift = lambda a: np.asarray(a)[None, None] if np.asarray(a).ndim == 3 else np.asarray(a)
    if average_across_subjects:
        per = [compute_metrics(lift(p), lift(l)) for p, l in zip(predictions, labels)]
        return {k: sum(m[k] for m in per) / len(per) for k in per[0]}
    return compute_metrics(np.concatenate([lift(p) for p in predictions]), np.concatenate([lift(l) for l in labels]))


def dice_iou(pred, label):
    """seg_fcd_test.py:160-178 for one subject ([B, C, ...] binary tensors): the empty-ground-truth edge case first, then
    DiceMetric / MeanIoU(include_background=False, reduction='mean')."""
    pred, label = np.asarray(pred), np.asarray(label)
    if label.sum() == 0:
        return (1.0, 1.0) if pred.sum() == 0 else (0.0, 0.0)
    c = _drop_background(confusion_counts(pred, label)).astype(np.float64)
    tp, fp, fn = c[..., 0], c[..., 1], c[..., 3]
    with np.errstate(divide="ignore", invalid="ignore"):
        dice = np.where(tp + fn > 0, 2 * tp / (2 * tp + fp + fn), np.nan)
        iou = np.where(tp + fn > 0, tp / (tp + fp + fn), np.nan)
    return _nanmean_channels_then_batch(dice), _nanmean_channels_then_batch(iou)


def evaluate_fp(cc, label):
    """utils/utils_common.py:37-60: number of component ids > 0 in `cc` that share no voxel with a non-zero label."""
    cc, label = np.asarray(cc), np.asarray(label)
    ids = np.unique(cc)
    ids = ids[ids > 0]
    hit = np.unique(cc[(label != 0) & (cc > 0)])
    return int(ids.size - hit.size)
