"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, the module trees mirror the
reference's state-dict layout, window enumeration matches the oracle, the loss parameter mapping, the no-CPU-fallback
guarantee, and the world_size-2 (gloo) data-parallel plumbing."""
import contextlib
import io
import os

import pytest
import torch

from tests import helpers as H


def test_library_loads_and_exports_every_declared_symbol():
    from fcd_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 30
    lib = _lib.lib()                       # ctypes.CDLL: works without a GPU
    for name in protos:
        assert hasattr(lib, name), f"libfcd_b200.so lacks {name} declared in include/fcd_b200.h"
    # argument-less sizing helpers run without touching the device state
    assert _lib.query("fcd_loss_blocks") > 0


def test_call_rejects_cpu_tensors_and_bad_arguments():
    from fcd_b200 import _lib
    with pytest.raises(TypeError):
        _lib.call("fcd_add", a=None)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        _lib._ptr(torch.zeros(4))


@pytest.mark.parametrize("name", H.MODEL_CASES)
def test_state_dict_layout_matches_reference(name):
    import fcd_b200
    meta, _ = H.load_case(name)
    params = fcd_b200.get_default_params()
    params.update(model_type=meta["model_type"], patch_size=(meta["patch"],) * 3, feature_size=meta["feature_size"])
    with contextlib.redirect_stdout(io.StringIO()):
        model, params = fcd_b200.get_model(params)
    sd = model.state_dict()
    assert list(sd.keys()) == [k for k, _, _ in meta["spec"]]
    for k, s, d in meta["spec"]:
        assert tuple(sd[k].shape) == tuple(s) and str(sd[k].dtype).endswith(d), k
    assert params["model_returns_vaeloss"] == ("vae" in meta["model_type"])


def test_default_config_parameter_counts():
    """43,524,802 (MS_DSA_NET) and 22,966,690 (BaseUNet) trainable parameters."""
    import fcd_b200
    for mt, n in (("MS_DSA_NET", 43524802), ("BaseUNet", 22966690)):
        params = fcd_b200.get_default_params()
        params.update(model_type=mt, patch_size=(128,) * 3)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            model, _ = fcd_b200.get_model(params)
        assert sum(p.numel() for p in model.parameters() if p.requires_grad) == n
        assert f"Trainable parameters: {n}" in buf.getvalue()      # get_model.py:246-248 prints it


def test_get_model_contract():
    import fcd_b200
    params = fcd_b200.get_default_params()
    params["model_type"] = "segresnetvae_dsa"
    model, p = fcd_b200.get_model(params, return_model=False)        # train.py:437
    assert model is None and p["model_returns_vaeloss"] is True
    params["model_type"] = "swinunetr"
    with pytest.raises(NotImplementedError):
        fcd_b200.get_model(params)


def test_no_cpu_fallback():
    import fcd_b200
    params = fcd_b200.get_default_params()
    params.update(model_type="baseunet", patch_size=(64,) * 3, feature_size=4)
    with contextlib.redirect_stdout(io.StringIO()):
        model, params = fcd_b200.get_model(params)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.zeros(1, 2, 64, 64, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fcd_b200.CombinedLoss(params, "cpu")(torch.zeros(1, 2, 8, 8, 8), torch.zeros(1, 1, 8, 8, 8))
    from fcd_b200.inferers import sliding_window_inference
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sliding_window_inference(torch.zeros(1, 2, 8, 8, 8), 8, 1, lambda x: x)


def test_product_never_imports_the_oracle():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fcd_b200")
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dp, f)


def test_loss_config_mapping():
    from fcd_b200.get_loss import CombinedLoss, loss_config
    p = H.loss_params(dict(loss_params=dict(loss="DiceFocalLoss", gamma_focal=3.0, lambda_focal=2.0,
                                            tv_loss_weight=0.1, tv_loss_norm="l2", tvloss_exclude_borders=True)))
    c = loss_config(p)
    assert c["kind"] == 2 and c["gamma"] == 3.0 and c["lambda_2"] == 2.0 and c["tv_norm"] == 2 and c["tv_exclude"] == 1
    g = loss_config(dict(p, loss="GeneralizedDiceFocalLoss", gdice_wtype="simple", lambda_dice=0.7))
    assert g["kind"] == 4 and g["w_type"] == 1 and g["lambda_dice"] == 0.7 and g["lambda_2"] == 2.0
    assert loss_config(dict(p, loss="GeneralizedDiceLoss"))["w_type"] == 0          # config.py default 'square'
    with pytest.raises(ValueError):
        loss_config(dict(p, loss="GeneralizedDiceLoss", gdice_wtype="cubic"))
    for bad in (dict(loss="TverskyLoss"), dict(sigmoid=True), dict(chans_out=3)):
        q = dict(p)
        q.update(bad)
        with pytest.raises(NotImplementedError):
            loss_config(q)
    q = dict(p)
    q["boundaryloss_weight"] = 0.3
    with pytest.raises(NotImplementedError):
        CombinedLoss(q, "cpu")


def test_window_enumeration_matches_oracle():
    from fcd_b200.inferers import window_starts
    from oracle import inferer as oinf
    for size in [(256, 256, 192), (182, 218, 182), (80, 72, 48), (24, 40, 32), (128, 128, 128), (130, 129, 257)]:
        for roi in (32, 128):
            for ov in (0.0, 0.25, 0.5, 0.75):
                r = (roi,) * 3
                s = tuple(max(a, b) for a, b in zip(size, r))
                assert window_starts(s, r, ov) == oinf.window_starts(s, r, ov)
    assert window_starts((256, 256, 192), (128,) * 3, 0.5) == [[0, 64, 128], [0, 64, 128], [0, 64]]   # 18 windows


def test_shard_range_partitions_work():
    from fcd_b200.parallel import shard_range
    for n in (1, 9, 18, 100):
        for w in (1, 2, 4, 8):
            got = sorted(i for r in range(w) for i in shard_range(n, r, w))
            assert got == list(range(n))


def _ddp_worker(rank, world, port, ret):
    import torch.distributed as dist
    from fcd_b200.parallel import GradAllReducer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(rank)                                   # replicas start DIFFERENT on purpose
        net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
        red = GradAllReducer(net.parameters())
        red.sync_params(src=0)
        w0 = [p.detach().clone() for p in net.parameters()]
        ws = [[torch.zeros_like(w) for _ in range(world)] for w in w0]
        for w, lst in zip(w0, ws):
            dist.all_gather(lst, w)
        same = all(torch.equal(lst[0], lst[1]) for lst in ws)

        def averaged_ok(reducer, scale):
            net.zero_grad(set_to_none=True)
            x = torch.full((4, 5), float(rank + 1) * scale)
            local_net = [p.detach().clone().requires_grad_(True) for p in net.parameters()]
            h = torch.nn.functional.linear(x, local_net[0], local_net[1])
            torch.nn.functional.linear(h, local_net[2], local_net[3]).sum().backward()
            local = [p.grad for p in local_net]                   # this rank's own gradients, untouched by hooks
            net(x).sum().backward()
            reducer.allreduce()
            gathered = [[torch.zeros_like(g) for _ in range(world)] for g in local]
            for g, lst in zip(local, gathered):
                dist.all_gather(lst, g)
            return all(torch.allclose(p.grad, sum(lst) / world, atol=1e-6) for p, lst in zip(net.parameters(), gathered))

        ok = averaged_ok(red, 1.0)
        # two buckets, the early one launched from inside backward: step 1 observes the readiness order, steps 2-3 overlap
        red2 = GradAllReducer(net.parameters(), overlap=True, early_fraction=0.3)
        ok = ok and averaged_ok(red2, 1.0)
        planned = red2.overlap and red2._early is not None and 0 < len(red2._early[0]) < 4
        ok = ok and planned and averaged_ok(red2, 2.0) and averaged_ok(red2, 3.0) and red2.early_launches == 2
        ret[rank] = bool(ok and same)
    finally:
        dist.destroy_process_group()


def test_grad_allreduce_world2_gloo():
    """N>1 data-parallel path on CPU: parameters broadcast from rank 0, gradients averaged over ranks."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ddp_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] is True and ret[1] is True


def test_window_shard_partitions_windows():
    """Sliding-window sharding: every window goes to exactly one rank, loads differ by at most one window."""
    from fcd_b200.inferers import window_shard
    for total in (1, 8, 18, 27):
        for world in (1, 2, 4, 8):
            parts = [window_shard(total, r, world) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(total))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_slab_bounds_and_label_assembly():
    """Sharded labels-only inference: the padded accumulation volume is cut into `world` equal D-slabs for the
    reduce-scatter; the all-gathered label slabs are put back together and the z padding is cropped."""
    from fcd_b200.inferers import assemble_label_slabs, slab_bounds
    for planes in (1, 7, 32, 182, 192, 256):
        for world in (1, 2, 3, 4, 8):
            b = [slab_bounds(planes, r, world) for r in range(world)]
            slab = b[0][2]
            assert all(x[2] == slab for x in b) and slab * world >= planes and (slab - 1) * world < planes
            assert b[0][0] == 0 and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
    g = torch.Generator().manual_seed(0)
    for (B, nch, Dp, H, W, pz, D, world) in [(1, 1, 32, 5, 6, 4, 24, 4), (2, 2, 37, 3, 4, 0, 37, 8), (1, 1, 16, 2, 2, 3, 9, 3)]:
        slab = slab_bounds(Dp, 0, world)[2]
        vol = torch.randint(0, 255, (B, nch, world * slab, H, W), generator=g, dtype=torch.uint8)
        full = torch.stack([vol[:, :, r * slab:(r + 1) * slab].permute(0, 2, 1, 3, 4) for r in range(world)])
        out = assemble_label_slabs(full, pz, D)
        assert torch.equal(out, vol[:, :, pz:pz + D])


def test_metric_reductions_from_count_tables():
    """Host side of fcd_b200.metrics (the ratios and MONAI's mean reductions on a [B, C, 4] confusion table; plain torch,
    so it runs without a GPU) against oracle/metrics.py: random tables, background dropped for C > 1, subjects without a
    lesion (Dice NaN, left out of the mean), all-empty tables (0 / 0 -> NaN precision and sensitivity, Dice 0)."""
    import math
    import numpy as np
    from fcd_b200 import metrics
    from oracle import metrics as om
    rng = np.random.default_rng(0)
    tables = [rng.integers(0, 1000, (B, C, 4)) for B, C in ((1, 1), (3, 1), (2, 2), (4, 3))]
    t = rng.integers(1, 50, (3, 1, 4))
    t[1, 0, [0, 3]] = 0                       # subject 1: no ground-truth voxel
    tables.append(t)
    tables.append(np.zeros((2, 1, 4), np.int64) + np.array([0, 0, 512, 0]))      # nothing predicted, nothing to find
    tables.append(np.array([[[0, 7, 100, 0]], [[0, 0, 107, 0]]]))                # false positives only
    for tab in tables:
        ref = om.metrics_from_counts(tab)
        got = metrics.metrics_from_counts(torch.from_numpy(np.asarray(tab, np.int64)))
        assert list(got) == ["Prec", "Sens", "F1", "DC"]
        for k in got:
            assert (math.isnan(got[k]) and math.isnan(ref[k])) or abs(got[k] - ref[k]) <= 1e-12 * max(1.0, abs(ref[k])), (k, tab)
        dev = metrics.metrics_from_counts(torch.from_numpy(np.asarray(tab, np.int64)), as_tensors=True)
        assert all(v.dim() == 0 and v.dtype == torch.float64 for v in dev.values())
    with pytest.raises(RuntimeError):
        metrics.confusion_counts(torch.zeros(1, 1, 2, 2, 2), torch.zeros(1, 1, 2, 2, 2))      # CPU tensors: no fallback
    with pytest.raises(RuntimeError):
        metrics.VoxelMetricAccumulator().aggregate()


def test_patch_sampler_host_logic():
    """GpuPatchSampler without a GPU: argument checks, MONAI's hole-size clipping, the mask cube edge of
    utils/gridmask.py:31, the probability ramp (get_transforms.py:116-126) and the no-CPU-fallback rule."""
    import math
    import fcd_b200
    s = fcd_b200.GpuPatchSampler(dict(patch_size=(20, 31, 27), samples_per_case=3), hole_size=(5, 40, 3))
    assert s.roi == (20, 31, 27) and s.hole_size == (5, 31, 3) and s.num_samples == 3
    assert s.hh == math.ceil(math.sqrt(20 * 20 + 31 * 31 + 27 * 27))
    assert not s.has_gradual_prob() and s.coarse_dropout_prob == 0.0 and s.gridmask_prob == 0.0
    s.set_prob(5, 10)
    assert s.coarse_dropout_prob == 0.0 and s.gridmask_prob == 0.0
    r = fcd_b200.GpuPatchSampler(dict(patch_size=16, coarse_dropout_max_prob=0.4, coarse_dropout_start_epoch=10,
                                      gridmask_max_prob=0.6, gridmask_start_epoch=0))
    assert r.has_gradual_prob() and r.gridmask_prob == 0.6 and r.coarse_dropout_prob == 0.0
    r.set_prob(5, 110)
    assert r.coarse_dropout_prob == 0.0 and r.gridmask_prob == pytest.approx(0.6 * 5 / 110)
    r.set_prob(60, 110)
    assert r.coarse_dropout_prob == pytest.approx(0.4 * 0.5) and r.gridmask_prob == pytest.approx(0.6 * 60 / 110)
    r.set_prob(500, 110)
    assert r.coarse_dropout_prob == pytest.approx(0.4) and r.gridmask_prob == pytest.approx(0.6)
    for bad in (dict(pos=0, neg=0), dict(holes=9), dict(grid_spacing_range=(8, 8))):
        with pytest.raises(ValueError):
            fcd_b200.GpuPatchSampler(dict(patch_size=16), **bad)
    with pytest.raises(RuntimeError):
        s(torch.zeros(2, 24, 32, 28), torch.zeros(1, 24, 32, 28), seed=1)
