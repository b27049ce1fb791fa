"""The validation loop of the reference on the device: ModelTrainer.inference (train.py:148-165) -> loss -> Activations(softmax)
+ AsDiscrete(threshold=0.5) -> ModelTrainer.post_process (train.py:167-182) -> voxel metrics (train.py:184-234,
metrics.py:74-160), one subject at a time, without the reference's per-subject host round trips (loss.item(), mask ->
numpy -> scipy -> tensor, whole volumes kept for torch.cat): sliding window, label map, connected components and the
confusion counts are CUDA kernels of this library; the only device->host reads are the final scalars."""
from __future__ import annotations

import torch

from . import metrics
from .inferers import post_process as _post_process
from .inferers import sliding_window_inference


def inference(model, inputs: torch.Tensor, params: dict, label_mode: str | None = None, sw_batch_size: int = 2,
              overlap: float = 0.25, **kw):
    """ModelTrainer.inference (train.py:148-165): sliding_window_inference(roi_size=params['patch_size'], sw_batch_size=2,
    overlap=0.25) over the model; tuple outputs (VAE models) are unwrapped as `_custom_predictor` does (train.py:150-154).
    The defaults are the reference's hard-coded values; a larger `sw_batch_size` runs fewer, fuller forwards (2.4x faster
    per 256 x 256 x 192 subject at 18).  Window order and the fp32 blend do not depend on it; the network's deep levels pick
    their split-K order from the row count, so the logits agree to bf16 rounding (tests/test_gpu_models.py::
    test_sliding_window_batching_is_invariant)."""
    predictor = model
    if not hasattr(model, "forward_cl"):
        def predictor(x):
            y = model(x)
            return y[0] if isinstance(y, (tuple, list)) else y
    return sliding_window_inference(inputs=inputs, roi_size=params["patch_size"], sw_batch_size=int(sw_batch_size),
                                    predictor=predictor, overlap=float(overlap), label_mode=label_mode, **kw)


@torch.no_grad()
def evaluate_subject(model, inputs: torch.Tensor, labels: torch.Tensor, params: dict, loss_function=None,
                     post_process: bool = True, sw_batch_size: int = 2, overlap: float = 0.25):
    """One iteration of the loop at train.py:196-215 (batch 1): returns (loss or None, fcd_prediction [D,H,W],
    fcd_label [D,H,W]) as device tensors.  The label map is `softmax >= 0.5` per channel, i.e. Activations(softmax=
    params['softmax']) + AsDiscrete(threshold=0.5) (train.py:185); sigmoid heads are not built (config.py: softmax)."""
    if params.get("sigmoid", False) or not params.get("softmax", True):
        raise NotImplementedError("evaluate: only the softmax head of the reference configuration is built")
    logits, lab = inference(model, inputs, params, label_mode="threshold", sw_batch_size=sw_batch_size, overlap=overlap)
    ch = 0 if logits.shape[1] == 1 else 1
    loss = loss_function(logits, labels) if loss_function is not None else None
    if post_process:
        lab = _post_process(lab, min_region_size=params.get("min_region_size", 50))
    return loss, lab[0, ch], labels[0, 0]


@torch.no_grad()
def evaluate(model, data_loader, params: dict, loss_function=None, device=None, post_process: bool = True,
             sw_batch_size: int = 2, overlap: float = 0.25):
    """ModelTrainer.evaluate (train.py:184-234) without the lesion-level / HD95 extras: returns (val_loss, metrics) with
    metrics = {'Prec', 'Sens', 'F1', 'DC'} (metrics.py:97-104, global over the subjects).  `data_loader` yields the
    reference's dictionaries {"image": [1,C,D,H,W], "label": [1,1,D,H,W]} (or (image, label) pairs)."""
    was_training = model.training
    model.eval()
    acc = metrics.VoxelMetricAccumulator()
    total, n = None, 0
    try:
        for item in data_loader:
            img, lab = (item["image"], item["label"]) if isinstance(item, dict) else item
            if device is not None:
                img, lab = img.to(device, dtype=torch.float32), lab.to(device, dtype=torch.float32)
            loss, pred, truth = evaluate_subject(model, img, lab, params, loss_function, post_process, sw_batch_size,
                                                 overlap)
            if loss is not None:
                total = loss.detach().float() if total is None else total + loss.detach().float()
            n += 1
            acc.update(pred, truth)
    finally:
        model.train(was_training)
    if n == 0:
        raise RuntimeError("evaluate: the data loader is empty")
    val_loss = float(total) / n if total is not None else float("nan")
    return val_loss, acc.aggregate()
