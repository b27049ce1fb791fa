"""TEST INFRASTRUCTURE ONLY -- imports the reference's own modules from /root/reference (read-only).

Used in THIS container to (a) validate oracle/nets.py, oracle/losses.py, oracle/inferer.py against the
reference's own source files and (b) generate tests/golden/*.  /root/reference does not exist on the GPU
box, so nothing in `-m gpu` tests, smoke() or bench.py may call this.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys

REFERENCE_ROOT = os.environ.get("FCD_REFERENCE_ROOT", "/root/reference")


def _load_by_path(name, path):
    """utils/__init__.py pulls monai.transforms (gridmask); utils_common.py itself needs numpy+scipy only."""
    import importlib.util
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "get_model.py"))


def load():
    """Return (get_model_module, get_loss_module, utils_common_module) of the reference."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    from . import monai_shim
    monai_shim.install()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    gm = importlib.import_module("get_model")
    gl = importlib.import_module("get_loss")
    uc = _load_by_path("fcd_ref_utils_common", os.path.join(REFERENCE_ROOT, "utils", "utils_common.py"))
    return gm, gl, uc


def load_gridmask():
    """utils/gridmask.py of the reference (numpy + torch; its MONAI base class comes from the shim)."""
    load()
    return _load_by_path("fcd_ref_gridmask", os.path.join(REFERENCE_ROOT, "utils", "gridmask.py"))


def load_transforms():
    """get_transforms.py of the reference (FCDTrainTransform: the probability ramp of coarse dropout / GridMask); its
    MONAI transform classes are inert records from the shim, its GridMaskd is the reference's own utils/gridmask.py."""
    load()
    return importlib.import_module("get_transforms")


def default_params():
    load()
    return importlib.import_module("config").get_default_params()


def build_model(params, seed=42, init_weights=True):
    """get_model(params) + model.apply(initialize_weights) as ModelTrainer.__init__ does (train.py:56-59)."""
    import torch
    gm, _, _ = load()
    tu = importlib.import_module("train_utils")
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        model, params = gm.get_model(params)
    if init_weights:
        model.apply(tu.initialize_weights)
    return model, params
