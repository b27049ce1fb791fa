"""TEST INFRASTRUCTURE ONLY -- CPU oracle for sliding-window inference and the label-map post-processing.

sliding_window_inference restates MONAI 1.5.1's mode='constant' path as called by the reference
(train.py:148-165, seg_fcd_test.py:37-54; SURVEY.md A7, parity UNPINNED against real MONAI).
post_process_segment restates utils/utils_common.py:10-33 (scipy.ndimage does the arithmetic there too).
label_map restates train.py:185,209-211 (softmax -> >= 0.5) and get_transforms.py:142-154 (argmax).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def window_starts(image_size, roi_size, overlap):
    """Per-dim window start lists: MONAI _get_scan_interval + dense_patch_slices (first dim slowest)."""
    starts = []
    for s, r in zip(image_size, roi_size):
        interval = r if r == s else max(int(r * (1 - overlap)), 1)
        num = int(math.ceil(float(s) / interval))
        scan = next((d for d in range(num) if d * interval + r >= s), None)
        n = scan + 1 if scan is not None else 1
        st = []
        for i in range(n):
            a = i * interval
            a -= max(a + r - s, 0)
            st.append(a)
        starts.append(st)
    return starts


def window_list(image_size, roi_size, overlap):
    sz, sy, sx = window_starts(image_size, roi_size, overlap)
    return [(z, y, x) for z in sz for y in sy for x in sx]


def sliding_window_inference(inputs, roi_size, sw_batch_size, predictor, overlap=0.25):
    """out[win] += pred (fp32, window order), cnt[win] += 1, out /= cnt; zero-pad images smaller than roi."""
    if isinstance(roi_size, int):
        roi_size = (roi_size,) * 3
    B = inputs.shape[0]
    orig = tuple(inputs.shape[2:])
    pads = []
    for k in (2, 1, 0):
        diff = max(roi_size[k] - orig[k], 0)
        pads.extend([diff // 2, diff - diff // 2])
    if any(pads):
        inputs = F.pad(inputs, pads)
    size = tuple(inputs.shape[2:])
    wins = window_list(size, roi_size, overlap)
    total = len(wins) * B
    out = cnt = None
    for g in range(0, total, sw_batch_size):
        idxs = list(range(g, min(g + sw_batch_size, total)))
        data = torch.cat([inputs[i // len(wins): i // len(wins) + 1, :,
                                 wins[i % len(wins)][0]: wins[i % len(wins)][0] + roi_size[0],
                                 wins[i % len(wins)][1]: wins[i % len(wins)][1] + roi_size[1],
                                 wins[i % len(wins)][2]: wins[i % len(wins)][2] + roi_size[2]] for i in idxs])
        pred = predictor(data)
        if isinstance(pred, (tuple, list)):
            pred = pred[0]
        if out is None:
            out = torch.zeros((B, pred.shape[1]) + size, dtype=inputs.dtype)
            cnt = torch.zeros((1, 1) + size, dtype=inputs.dtype)
            for (z, y, x) in wins:
                cnt[:, :, z:z + roi_size[0], y:y + roi_size[1], x:x + roi_size[2]] += 1
        for j, i in enumerate(idxs):
            z, y, x = wins[i % len(wins)]
            b = i // len(wins)
            out[b:b + 1, :, z:z + roi_size[0], y:y + roi_size[1], x:x + roi_size[2]] += pred[j:j + 1].to(out.dtype)
    out = out / cnt
    if any(pads):
        out = out[:, :, pads[4]:pads[4] + orig[0], pads[2]:pads[2] + orig[1], pads[0]:pads[0] + orig[2]]
    return out


def label_map(logits, mode="threshold"):
    """'threshold': Activations(softmax) + AsDiscrete(threshold=0.5) -> [B,C,...] float {0,1}; 'argmax': [B,1,...]."""
    if mode == "argmax":
        return torch.argmax(logits, dim=1, keepdim=True)
    return (torch.softmax(logits.float(), dim=1) >= 0.5).float()


def post_process_segment(mask, l_min):
    """utils/utils_common.py:10-33: opening (6-conn) -> fill holes (5^3) -> 26-conn label -> size filter."""
    from scipy import ndimage as nd
    out_msk = np.zeros_like(mask)
    out_lab = np.zeros_like(mask)
    morphed = nd.binary_opening(mask, iterations=1)
    morphed = nd.binary_fill_holes(morphed, structure=np.ones((5, 5, 5))).astype(int)
    lab, _ = nd.label(morphed, structure=np.ones((3, 3, 3)))
    vals = np.unique(lab)
    sizes = nd.labeled_comprehension(morphed, lab, vals, np.sum, float, 0)
    if l_min == -1:
        l_min = np.max(sizes)
    count = 0
    for l in range(len(sizes)):
        if sizes[l] >= l_min:
            count += 1
            sel = lab == l
            out_msk[sel] = 1
            out_lab[sel] = count
    return out_msk, out_lab
