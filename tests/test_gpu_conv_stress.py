"""Persistent-CTA regimes of the tcgen05 conv kernels that the small parity cases never reach: several work items per
CTA (nitems > 2 x 296), several d-segments per column, and both at once -- the regime in which round 1 saw rare
bounded-wait time-outs inside 4-5 window inference batches.  Every case is called through the C ABI with the segment
count FORCED (the host chooser would not pick all of them), compared with F.conv3d in fp32 once, and then launched
repeatedly with a second stream keeping other tcgen05 convs (512 TMEM columns, one CTA per SM) resident on the same
SMs; every launch must reproduce the first result bit for bit (the kernels are deterministic) and leave the status
word (include/fcd_b200.h, fcd_status) at zero."""
import os
import subprocess
import sys

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ops():
    from fcd_b200 import ops as _ops
    return _ops


def _conv(ops, entry, x, w32, Ci, Co, nseg, flip=0, stats=True):
    B, D, H, W, Kp = x.shape
    Np = ops.pad16(Co)
    y = torch.empty((B, D, H, W, Np), dtype=torch.bfloat16, device=x.device)
    part = None
    fin = dict(ops._NOFIN)
    if stats and Np <= 32:
        part = torch.empty((B, ((H + 15) // 16) * ((W + 7) // 8) * nseg, 2, Np), dtype=torch.float32, device=x.device)
        from fcd_b200 import _lib
        if _lib.lib().fcd_norm_fin_fold(B, part.shape[1], 2 * Np):
            # InstanceNorm statistics finished by the conv's last CTA
            fin.update(mean=torch.empty((B, Np), dtype=torch.float32, device=x.device),
                       rstd=torch.empty((B, Np), dtype=torch.float32, device=x.device), norm_mode=0, eps=1e-5)
    if flip:    # data gradient: x plays dY (Ci = conv's Cout), the weight is read transposed and mirrored
        ops._conv3_call(entry, A=x, lda=Kp, Wf=w32, Nr=Co, Kr=Ci, sn=27, sk=Co * 27, st=1, kseg=Ci, ksegpad=Kp, nsg=Co,
                        nsgpad=Np, C=y, ldc=Np, part=None, Bn=B, D=D, H=H, W=W, K=Kp, N=Np, flip=1, nseg=nseg, **ops._NOFIN)
        part = None
    else:
        ops._conv3_call(entry, A=x, lda=Kp, Wf=w32, Nr=Co, Kr=Ci, sn=Ci * 27, sk=27, st=1, kseg=Ci, ksegpad=Kp, nsg=Co,
                        nsgpad=Np, C=y, ldc=Np, part=part, Bn=B, D=D, H=H, W=W, K=Kp, N=Np, flip=0, nseg=nseg, **fin)
        if part is not None and fin["mean"] is not None:
            part = torch.cat([part.reshape(B, -1), fin["mean"], fin["rstd"]], 1)    # compared bit for bit across launches
            y._mean_rstd = (fin["mean"], fin["rstd"])
    return y, part


# B, Ci, Co, S (cube edge), nseg, flip, iterations
STRESS_CASES = [
    (5, 16, 16, 128, 1, 0, 60),
    (5, 16, 16, 128, 2, 0, 300),      # 1280 items on 296 CTAs: segments AND several items per CTA
    (5, 32, 16, 128, 1, 0, 40),
    (5, 32, 16, 128, 2, 0, 300),
    (5, 16, 16, 128, 2, 1, 100),      # data-gradient orientation of the same
    (5, 32, 32, 64, 8, 0, 300),       # 160 columns x 8 segments of 8 planes: 2-9 SHORT items per CTA
    (4, 64, 32, 64, 8, 0, 600),       # decoder2.conv1 of a 4-window batch: the shape round 2's reproducer timed out in
    (5, 64, 32, 64, 8, 0, 300),
    (5, 16, 32, 64, 8, 0, 200),
    (5, 64, 32, 32, 8, 0, 300),       # 40 columns x 8 segments of FOUR planes, 4-stage ring (DEPTH 1), one CTA per SM
    (5, 64, 32, 32, 8, 1, 100),
    (5, 32, 32, 32, 8, 0, 200),
    (9, 16, 16, 64, 16, 0, 200),      # 4-plane items, two CTAs per SM
    (3, 64, 16, 64, 4, 0, 100),
]


@pytest.mark.parametrize("B,Ci,Co,S,nseg,flip,iters", STRESS_CASES)
def test_tcf_many_items_and_segments(ops, B, Ci, Co, S, nseg, flip, iters):
    from fcd_b200 import _lib
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(S + 7 * nseg + Ci)
    Kp = ops.pad16(Ci)
    x = (torch.randn((B, S, S, S, Kp), generator=g) * 0.5).to(torch.bfloat16).to(dev)
    if flip:
        w = (torch.randn((Ci, Co, 3, 3, 3), generator=g) * 0.05)        # conv weight [Cout = Ci][Cin = Co]
        ref = F.conv_transpose3d(x.float().permute(0, 4, 1, 2, 3), w.to(dev), padding=1)
    else:
        w = (torch.randn((Co, Ci, 3, 3, 3), generator=g) * 0.05)
        ref = F.conv3d(x.float().permute(0, 4, 1, 2, 3), w.to(dev), padding=1)
    w32 = w.to(dev).contiguous()
    nitems = B * ((S + 15) // 16) * ((S + 7) // 8) * nseg
    _lib.status()       # clear
    y0, p0 = _conv(ops, "fcd_conv3_tcf", x, w32, Ci, Co, nseg, flip)
    _lib.check_errors()
    got = y0.float().permute(0, 4, 1, 2, 3)
    rel = float((got - ref).norm() / ref.norm())
    assert rel <= 4e-3, f"fcd_conv3_tcf vs F.conv3d: rel L2 {rel:.3e}"
    del ref, got
    if p0 is not None:
        yf = y0.float().reshape(B, -1, y0.shape[-1])
        if hasattr(y0, "_mean_rstd"):
            mean, rstd = y0._mean_rstd
            assert torch.allclose(mean, yf.mean(1), rtol=1e-4, atol=1e-5), "fused mean (finished by the last CTA)"
            assert torch.allclose(rstd[:, :Co], torch.rsqrt(yf.var(1, unbiased=False) + 1e-5)[:, :Co], rtol=2e-4)
        else:
            assert torch.allclose(p0.sum(1)[:, 0], yf.sum(1), rtol=1e-4, atol=1e-2)
        del yf
    # other tcgen05 convs resident on the same SMs (what the branch streams of a window forward do)
    side = torch.cuda.Stream()
    x2 = (torch.randn((2, 64, 64, 64, 32), generator=g) * 0.5).to(torch.bfloat16).to(dev)
    w2 = (torch.randn((32, 32, 3, 3, 3), generator=g) * 0.05).to(dev).contiguous()
    torch.cuda.synchronize()
    bad = 0
    for i in range(iters):
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                _conv(ops, "fcd_conv3_tcf", x2, w2, 32, 32, 1)
        y, p = _conv(ops, "fcd_conv3_tcf", x, w32, Ci, Co, nseg, flip)
        torch.cuda.current_stream().wait_stream(side)
        same = torch.equal(y, y0) and (p is None or torch.equal(p, p0))
        if not same:
            bad += 1
    st = _lib.status()
    assert st["word"] == 0, f"bounded wait timed out: {st} ({nitems} items)"
    assert bad == 0, f"{bad} of {iters} launches differ from the first ({nitems} items)"


@pytest.mark.parametrize("B,Ci,Co,S,nseg", [(5, 16, 64, 64, 4), (5, 32, 64, 32, 4), (3, 16, 16, 128, 2)])
def test_plain_tc_many_items_and_segments(ops, B, Ci, Co, S, nseg):
    """The same regimes for the plain tcgen05 kernel (Cout = 64 layers, and the fall-back of the kd-folded one)."""
    from fcd_b200 import _lib
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(11)
    x = (torch.randn((B, S, S, S, ops.pad16(Ci)), generator=g) * 0.5).to(torch.bfloat16).to(dev)
    w = (torch.randn((Co, Ci, 3, 3, 3), generator=g) * 0.05).to(dev).contiguous()
    ref = F.conv3d(x.float().permute(0, 4, 1, 2, 3), w, padding=1)
    _lib.status()
    y0, _ = _conv(ops, "fcd_conv3_tc", x, w, Ci, Co, nseg)
    rel = float((y0.float().permute(0, 4, 1, 2, 3) - ref).norm() / ref.norm())
    assert rel <= 4e-3, rel
    del ref
    for _ in range(100):
        y, _ = _conv(ops, "fcd_conv3_tc", x, w, Ci, Co, nseg)
        assert torch.equal(y, y0)
    _lib.check_errors()


def test_slow_producers_cannot_make_the_mma_warps_read_a_stale_plane():
    """Regression test of round 1's defect (root-caused in round 2, see the FULL wait in csrc/conv_tcf.cu): with the
    producers slowed down (FCD_TCF_PRODUCER_DELAY_NS) the MMA warps are always waiting for the next halo plane, which is
    when a warp that jumped 5 ring positions over an item boundary with one padding plane used to see the parity of
    plane sq - 2 NST and multiply a stale plane (4 of 4 runs failed with time-outs: profiles/r02_defect_ab.txt).  The
    4-stage-ring shapes run in a subprocess because the switch is read once per process."""
    env = dict(os.environ, FCD_TCF_PRODUCER_DELAY_NS="3000")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_conv_stress.py"), "-q", "-m",
                        "gpu", "-p", "no:cacheprovider", "-k", "test_tcf_many and (64-32-64-8 or 64-32-32-8)"],
                       capture_output=True, text=True, env=env, cwd=ROOT, timeout=1200)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0 and "4 passed" in r.stdout


def test_window_forward_graph_replays_with_unrestricted_segments():
    """The regime round 1's defect was seen in: CUDA-graph replays of the whole MS_DSA_NET forward on a 5-window batch
    (128^3 levels get 2 segments, 64^3 / 32^3 levels 8 short ones; branch streams active); every replay must reproduce
    the first bit for bit with a clean status word.  Subprocess: a fresh CUDA context and graph pool."""
    env = dict(os.environ)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress_window_forward.py"), "300", "5"],
                       capture_output=True, text=True, env=env, cwd=ROOT, timeout=1500)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0, "window-forward stress failed (see output)"
    assert "RESULT ok" in r.stdout
