"""CPU model of the deduplicated union rules of csrc/ccl.cu `uf_merge` (the kernel's comment cites it): on random small
volumes the rules must produce the partition of (a) the full forward walk they replace and (b) scipy's 26-connected
labelling for the foreground pass.  Pure Python loops, so sizes are tiny; row lengths straddle the 32-voxel segments the
device's run initialisation works in.  The device kernel itself is checked bit-exact against scipy in
tests/test_gpu_inference.py."""
import numpy as np
import pytest
from scipy import ndimage as nd


def _find(L, a):
    while L[a] != a:
        L[a] = L[L[a]]
        a = L[a]
    return a


def _union(L, a, b):
    a, b = _find(L, a), _find(L, b)
    if a < b:
        L[b] = a
    elif b < a:
        L[a] = b


def partition(active, near, dedup):
    """Roots after uf_init + uf_merge.  active[z,y,x]; near[z,y,x]: reach 2 (background voxels next to the foreground)."""
    D, H, W = active.shape
    N = D * H * W
    L = np.full(N, -1, np.int64)
    idx = lambda z, y, x: (z * H + y) * W + x
    for z in range(D):                               # uf_init: x-runs inside 32-voxel segments of the linear index
        for y in range(H):
            for x in range(W):
                if active[z, y, x]:
                    i = idx(z, y, x)
                    L[i] = L[i - 1] if (x > 0 and (i & 31) != 0 and active[z, y, x - 1]) else i
    unions = 0
    for z in range(D):
        for y in range(H):
            for x in range(W):
                if not active[z, y, x]:
                    continue
                i = idx(z, y, x)
                R = 2 if near[z, y, x] else 1
                joined = x > 0 and (i & 31) != 0 and active[z, y, x - 1]
                Rp = (2 if near[z, y, x - 1] else 1) if joined else 0
                for dx in range(1, R + 1):           # my own row
                    if x + dx < W and active[z, y, x + dx] and not (dx == 1 and (i & 31) != 31):
                        _union(L, i, i + dx)
                        unions += 1
                for dz in range(0, R + 1):
                    for dy in range(-R, R + 1):
                        if dz == 0 and dy <= 0:
                            continue
                        zz, yy = z + dz, y + dy
                        if zz >= D or yy < 0 or yy >= H:
                            continue
                        jrow = idx(zz, yy, 0)
                        if not dedup:                # the walk the rules replace: every forward neighbour
                            for xx in range(max(x - R, 0), min(x + R, W - 1) + 1):
                                if active[zz, yy, xx]:
                                    _union(L, i, jrow + xx)
                                    unions += 1
                            continue
                        covered = Rp >= max(dz, abs(dy))
                        lo = x + Rp if covered else x - R
                        prev_act = bool(covered and lo - 1 < W and active[zz, yy, lo - 1])
                        for xx in range(lo, x + R + 1):
                            if xx < 0:
                                continue
                            if xx >= W:
                                break
                            a = bool(active[zz, yy, xx])
                            if a and not (prev_act and ((jrow + xx) & 31) != 0):
                                _union(L, i, jrow + xx)
                                unions += 1
                            prev_act = a
    return np.array([_find(L, i) if L[i] >= 0 else -1 for i in range(N)]).reshape(D, H, W), unions


@pytest.mark.parametrize("seed", range(4))
def test_deduplicated_unions_give_the_same_components(seed):
    rng = np.random.default_rng(seed)
    saved = []
    for _ in range(5):
        D, H, W = int(rng.integers(3, 6)), int(rng.integers(3, 7)), int(rng.choice([5, 17, 31, 32, 33, 40]))
        fg = rng.random((D, H, W)) < rng.choice([0.15, 0.4, 0.6, 0.85])
        none = np.zeros_like(fg)
        full, n_full = partition(fg, none, False)
        ded, n_ded = partition(fg, none, True)
        assert np.array_equal(full, ded)
        # scipy numbers 26-connected components in raster order of their first voxel = ascending root index
        lab, n = nd.label(fg, structure=np.ones((3, 3, 3)))
        roots = np.unique(ded[ded >= 0])
        assert len(roots) == n
        assert np.array_equal(np.searchsorted(roots, ded[fg]) + 1, lab[fg])
        # background pass of binary_fill_holes(structure = ones 5^3): reach 2 next to the foreground
        bg = ~fg
        near = nd.binary_dilation(fg, structure=np.ones((3, 3, 3), bool)) & bg
        full_b, m_full = partition(bg, near, False)
        ded_b, m_ded = partition(bg, near, True)
        assert np.array_equal(full_b, ded_b)
        saved.append((n_full, n_ded, m_full, m_ded))
    assert sum(s[1] for s in saved) < sum(s[0] for s in saved) and sum(s[3] for s in saved) < sum(s[2] for s in saved)
