"""TEST INFRASTRUCTURE ONLY -- portable, bit-reproducible synthetic weights / inputs / labels.

torch's RNG streams are not guaranteed identical across builds, so goldens are keyed on a counter-based
integer hash (splitmix64 in numpy uint64 arithmetic): the same (name, shape, seed) gives the same bits in
this container and on the GPU box.  All values are rounded to bf16-representable fp32 so that the CUDA
path (bf16 operands, fp32 accumulate) and the fp32 oracle consume identical numbers.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def uniform(n: int, key: str, seed: int = 0) -> np.ndarray:
    """n float64 uniforms in [0,1), keyed by (key, seed)."""
    base = np.uint64((zlib.crc32(key.encode()) << 20) ^ (seed * 0x51ED27))
    idx = np.arange(n, dtype=np.uint64) + base
    return (_splitmix64(idx) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def normalish(n: int, key: str, seed: int = 0) -> np.ndarray:
    """Approximately N(0,1): sum of 4 uniforms, centred and scaled (portable, no transcendental functions)."""
    u = sum(uniform(n, f"{key}#{i}", seed) for i in range(4))
    return (u - 2.0) * np.sqrt(3.0)


def to_bf16_exact(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def tensor(shape, key, seed=0, scale=1.0, shift=0.0, dist="uniform") -> torch.Tensor:
    n = int(np.prod(shape)) if len(shape) else 1
    v = (uniform(n, key, seed) * 2.0 - 1.0) if dist == "uniform" else normalish(n, key, seed)
    t = torch.from_numpy((v * scale + shift).astype(np.float32)).reshape(tuple(shape))
    return to_bf16_exact(t)


def synthetic_state_dict(spec, seed=0):
    """spec: iterable of (name, shape, dtype_str).  Returns {name: tensor} with per-kind scales.

    Scales follow the reference's initialisers in spirit (train_utils.py:44-60 kaiming fan_out for convs,
    conv_blocks.py:145-149 for EF) but, unlike the reference's gamma=1e-6 / pos_embed=0, keep every branch
    numerically visible so parity tests exercise DSA, pos_embed and the affine norms.
    """
    sd = {}
    for name, shape, dtype in spec:
        shape = tuple(shape)
        leaf = name.split(".")[-1]
        if leaf == "num_batches_tracked":
            sd[name] = torch.zeros(shape, dtype=torch.int64)
        elif leaf == "running_mean":
            sd[name] = tensor(shape, name, seed, 0.1)
        elif leaf == "running_var":
            sd[name] = tensor(shape, name, seed, 0.2, 1.0)
        elif leaf == "pos_embed":
            sd[name] = tensor(shape, name, seed, 0.1)
        elif leaf == "gamma":
            sd[name] = tensor(shape, name, seed, 0.25, 0.5)
        elif leaf in ("temperature", "temperature2"):
            sd[name] = tensor(shape, name, seed, 0.3, 1.0)
        elif leaf == "EF":
            sd[name] = tensor(shape, name, seed, 1.0 / np.sqrt(shape[-1]))
        elif leaf == "bias":
            sd[name] = tensor(shape, name, seed, 0.1)
        elif leaf == "weight" and len(shape) == 1:                       # LN / BN / GN affine
            sd[name] = tensor(shape, name, seed, 0.2, 1.0)
        elif leaf == "weight" and len(shape) == 5:                       # Conv3d [Co,Ci,k,k,k] / ConvTranspose3d
            fan = shape[0] * shape[2] * shape[3] * shape[4]
            if "transp_conv" in name or ".deconv." in name:
                fan = shape[1] * shape[2] * shape[3] * shape[4] / 8.0
            sd[name] = tensor(shape, name, seed, np.sqrt(2.0 / fan) * np.sqrt(3.0))
        elif leaf == "weight" and len(shape) == 2:                       # Linear
            sd[name] = tensor(shape, name, seed, np.sqrt(6.0 / (shape[0] + shape[1])))
        else:
            sd[name] = tensor(shape, name, seed, 0.1)
    return sd


def spec_of(state_dict):
    return [(k, tuple(v.shape), str(v.dtype).replace("torch.", "")) for k, v in state_dict.items()]


def image(batch, chans, size, seed=0):
    """Synthetic 2-channel MRI-like patch [B,C,D,H,W], bf16-exact, roughly unit variance."""
    if isinstance(size, int):
        size = (size,) * 3
    return tensor((batch, chans) + tuple(size), "image", seed, 1.0, 0.0, dist="normal")


def label(batch, size, seed=0, n_blobs=3):
    """Binary lesion mask [B,1,D,H,W] float {0,1}: a few ellipsoids (~0.5-2 % foreground)."""
    if isinstance(size, int):
        size = (size,) * 3
    D, H, W = size
    zz, yy, xx = np.meshgrid(np.arange(D), np.arange(H), np.arange(W), indexing="ij")
    out = np.zeros((batch, 1, D, H, W), dtype=np.float32)
    for b in range(batch):
        u = uniform(6 * n_blobs, f"label{b}", seed).reshape(n_blobs, 6)
        for k in range(n_blobs):
            c = u[k, :3] * np.array([D, H, W]) * 0.7 + np.array([D, H, W]) * 0.15
            r = (u[k, 3:] * 0.08 + 0.05) * np.array([D, H, W])
            m = ((zz - c[0]) / r[0]) ** 2 + ((yy - c[1]) / r[1]) ** 2 + ((xx - c[2]) / r[2]) ** 2 <= 1.0
            out[b, 0][m] = 1.0
    return torch.from_numpy(out)
