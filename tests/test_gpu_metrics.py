"""Voxel-level evaluation on the device (fcd_b200/metrics.py, csrc/metrics.cu) against the numpy oracle
(oracle/metrics.py, which restates metrics.py:74-160 and utils/utils_common.py:37-60): the confusion counts and the
false-positive component count BIT-EXACT (integer work), the ratios to 1e-12 (a few fp64 scalar operations)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _pair(shape, seed, p_fg=0.3, t_fg=0.2):
    g = np.random.default_rng(seed)
    pred = (g.random(shape) < p_fg).astype(np.float32)
    lab = (g.random(shape) < t_fg).astype(np.float32)
    return pred, lab


def _close(a, b):
    return (math.isnan(a) and math.isnan(b)) or abs(a - b) <= 1e-12 * max(1.0, abs(b))


@pytest.mark.parametrize("shape", [(1, 1, 40, 48, 56), (3, 2, 17, 19, 23), (2, 3, 8, 8, 9), (1, 2, 64, 64, 64),
                                   (2, 1, 1, 1, 5)])
def test_confusion_counts_bit_exact(shape):
    from fcd_b200 import metrics
    from oracle import metrics as om
    pred, lab = _pair(shape, sum(shape))
    soft = pred * 0.4 + 0.3 + np.random.default_rng(1).random(shape, dtype=np.float32) * 0.05     # values around 0.5
    for p in (pred, soft):
        got = metrics.confusion_counts(torch.from_numpy(p).to(DEV), torch.from_numpy(lab).to(DEV)).cpu().numpy()
        ref = om.confusion_counts(p, lab)
        assert np.array_equal(got, ref)
        assert np.all(got.sum(-1) == np.prod(shape[2:]))
    # uint8 / bool predictions (the argmax label map of sliding_window_inference), unaligned views
    got = metrics.confusion_counts(torch.from_numpy(pred).to(DEV).to(torch.uint8), torch.from_numpy(lab).to(DEV))
    assert np.array_equal(got.cpu().numpy(), om.confusion_counts(pred, lab, 0.5))
    got = metrics.confusion_counts(torch.from_numpy(pred).to(DEV) > 0, torch.from_numpy(lab).to(DEV).long())
    assert np.array_equal(got.cpu().numpy(), om.confusion_counts(pred, lab, 0.5))


def test_voxel_level_metrics_follow_the_reference_reductions():
    from fcd_b200 import metrics
    from oracle import metrics as om
    subjects = [_pair((24, 28, 20), s, 0.1 + 0.1 * s, 0.15) for s in range(4)]
    subjects.append((subjects[0][0], np.zeros_like(subjects[0][1])))          # a subject without a lesion: Dice NaN, left out
    preds = [torch.from_numpy(p).to(DEV) for p, _ in subjects]
    labs = [torch.from_numpy(l).to(DEV) for _, l in subjects]
    for avg in (False, True):
        got = metrics.calculate_voxel_level_metrics(preds, labs, average_across_subjects=avg)
        ref = om.calculate_voxel_level_metrics([p for p, _ in subjects], [l for _, l in subjects], avg)
        assert list(got) == ["Prec", "Sens", "F1", "DC"]
        for k in got:
            assert _close(got[k], ref[k]), (avg, k, got[k], ref[k])
    # two-channel tensors: the background channel is dropped
    p2 = np.stack([1 - subjects[1][0], subjects[1][0]])[None]
    l2 = np.stack([1 - subjects[1][1], subjects[1][1]])[None]
    got = metrics.compute_metrics(torch.from_numpy(p2).to(DEV), torch.from_numpy(l2).to(DEV))
    ref = om.compute_metrics(p2, l2)
    for k in got:
        assert _close(got[k], ref[k])
    one = om.compute_metrics(subjects[1][0][None, None], subjects[1][1][None, None])
    assert all(_close(ref[k], one[k]) for k in got)
    # nothing predicted, nothing to find: precision / sensitivity are NaN (0 / 0), Dice 0 (every subject left out)
    z = torch.zeros((1, 1, 8, 8, 8), device=DEV)
    got = metrics.compute_metrics(z, z)
    assert math.isnan(got["Prec"]) and math.isnan(got["Sens"]) and got["DC"] == 0.0
    with pytest.raises(NotImplementedError):
        metrics.compute_metrics(z, z, compute_hd95=True)


def test_dice_iou_and_the_empty_ground_truth_edge_case():
    from fcd_b200 import metrics
    from oracle import metrics as om
    pred, lab = _pair((1, 2, 20, 24, 28), 5)
    cases = [(pred, lab), (pred, np.zeros_like(lab)), (np.zeros_like(pred), np.zeros_like(lab)),
             (np.zeros_like(pred), lab)]
    for p, l in cases:
        got = metrics.dice_iou(torch.from_numpy(p).to(DEV), torch.from_numpy(l).to(DEV))
        ref = om.dice_iou(p, l)
        assert _close(got[0], ref[0]) and _close(got[1], ref[1]), (got, ref)


def test_evaluate_fp_counts_components_without_overlap():
    import fcd_b200
    from fcd_b200 import metrics
    from oracle import metrics as om
    g = np.random.default_rng(11)
    D, H, W = 48, 56, 40
    mask = np.zeros((D, H, W), np.float32)
    for _ in range(14):
        c = [int(g.integers(5, s - 5)) for s in (D, H, W)]
        mask[c[0] - 3:c[0] + 3, c[1] - 3:c[1] + 3, c[2] - 3:c[2] + 3] = 1
    lab = np.zeros_like(mask)
    lab[:24] = mask[:24] * (g.random((24, H, W)) < 0.5)                   # truth overlaps some components only
    _, cc = fcd_b200.post_process_segment(torch.from_numpy(mask).to(DEV), 5)
    got = int(metrics.evaluate_fp(cc, torch.from_numpy(lab).to(DEV)))
    ref = om.evaluate_fp(cc.cpu().numpy(), lab)
    assert got == ref and 0 < ref < int(cc.max())
    # arbitrary (non-compact) ids, an empty component volume, and an id beyond max_id
    ids = (g.integers(0, 300, (D, H, W)) * (g.random((D, H, W)) < 0.01)).astype(np.float32)
    got = int(metrics.evaluate_fp(torch.from_numpy(ids).to(DEV), torch.from_numpy(lab).to(DEV), max_id=299))
    assert got == om.evaluate_fp(ids, lab)
    assert int(metrics.evaluate_fp(torch.zeros((D, H, W), device=DEV), torch.from_numpy(lab).to(DEV))) == 0
    assert int(metrics.evaluate_fp(torch.from_numpy(ids).to(DEV), torch.from_numpy(lab).to(DEV), max_id=10)) == -1


def test_evaluate_loop_matches_the_oracle_pipeline():
    """fcd_b200.evaluate (ModelTrainer.evaluate, train.py:184-234) on three synthetic subjects of different sizes with an
    exact-arithmetic predictor: the oracle pipeline is oracle sliding window (roi = patch_size, sw_batch_size 2, overlap
    0.25) -> softmax >= 0.5 -> scipy post_process_segment(min_region_size) -> numpy confusion counts / metrics; the
    post-processed masks and count tables must be identical, the metrics agree to 1e-12, the loss to fp32 accuracy."""
    import fcd_b200
    from fcd_b200 import evaluation, metrics
    from oracle import inferer as oinf
    from oracle import losses as olosses
    from oracle import metrics as om
    from oracle import synth
    from tests.test_gpu_inference import _ExactNet, _dyadic

    params = fcd_b200.get_default_params()
    params.update(patch_size=(32, 32, 32), min_region_size=20, loss="DiceCELoss")
    w = _dyadic((2, 2, 3, 3, 3), 5, 0.125, 0.5)
    net = _ExactNet(w)
    net.train()
    sizes = [(48, 40, 56), (32, 64, 40), (40, 40, 40)]
    data, ref_tables, ref_losses, ref_masks = [], [], [], []
    for i, size in enumerate(sizes):
        x = _dyadic((1, 2) + size, 10 + i, 0.25, 2.0)
        y = synth.label(1, size, seed=30 + i, n_blobs=3)
        data.append({"image": x, "label": y})
        logits = oinf.sliding_window_inference(x, params["patch_size"], 2, net.forward, 0.25)
        ref_losses.append(float(olosses.combined_loss(params, logits, y)))
        lab = oinf.label_map(logits, "threshold")
        mask, _ = oinf.post_process_segment((lab[0, 1] > 0.5).float().numpy(), params["min_region_size"])
        ref_masks.append(mask)
        ref_tables.append(om.confusion_counts(mask[None, None], y.numpy()))
    loss_fn = fcd_b200.CombinedLoss(params, DEV)
    # per subject: masks and count tables
    acc = metrics.VoxelMetricAccumulator()
    for d, mask, tab in zip(data, ref_masks, ref_tables):
        loss, pred, truth = evaluation.evaluate_subject(net, d["image"].to(DEV), d["label"].to(DEV), params, loss_fn)
        assert np.array_equal(pred.cpu().numpy(), mask)
        acc.update(pred, truth)
        assert np.array_equal(acc.tables[-1].cpu().numpy(), tab)
    # the whole loop (subjects of different sizes: the reference's torch.cat could not even take them)
    val_loss, got = fcd_b200.evaluate(net, data, params, loss_fn, device=DEV)
    assert net.training                                           # the mode is restored
    ref = om.metrics_from_counts(np.concatenate(ref_tables))
    for k in ("Prec", "Sens", "F1", "DC"):
        assert _close(got[k], ref[k]), (k, got[k], ref[k])
    assert abs(val_loss - sum(ref_losses) / 3) <= 2e-5 * max(1.0, abs(sum(ref_losses) / 3))
    # without post-processing the thresholded FCD channel itself is scored
    _, pred, _ = evaluation.evaluate_subject(net, data[0]["image"].to(DEV), data[0]["label"].to(DEV), params, None, False)
    logits = oinf.sliding_window_inference(data[0]["image"], params["patch_size"], 2, net.forward, 0.25)
    assert np.array_equal(pred.cpu().numpy(), oinf.label_map(logits, "threshold")[0, 1].numpy())


def test_evaluate_fp_matches_the_reference_goldens():
    """tests/golden/evaluate_fp_cases.json holds utils/utils_common.py:37-60's own counts (tools/make_goldens.py)."""
    import json
    import os
    from scipy import ndimage as nd
    from fcd_b200 import metrics
    from oracle import synth
    from tests import helpers as H
    for c in json.load(open(os.path.join(H.GOLDEN, "evaluate_fp_cases.json"))):
        k = c["seed"]
        pred = nd.binary_dilation(synth.tensor((24, 28, 20), "fp_pred", 40 + k, 1.0, dist="normal").numpy() > 2.6,
                                  iterations=1 + k % 3)
        lab = nd.binary_dilation(synth.tensor((24, 28, 20), "fp_lab", 50 + k, 1.0, dist="normal").numpy() > 2.0,
                                 iterations=2).astype(np.float32)
        cc, n = nd.label(pred)
        got = int(metrics.evaluate_fp(torch.from_numpy(cc.astype(np.float32)).to(DEV), torch.from_numpy(lab).to(DEV)))
        assert n == c["components"] and got == c["fp"], c
