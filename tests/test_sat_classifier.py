"""CPU: the three-way SAT classifier (obj2voxel_b200/csrc/o2v_sat.cuh) that decides which (leaf, voxel) pairs the kernels
send to the exact clip, fuzzed against the reference semantics (plane-distance cull + exact six-plane clip, the host
build of o2v_exact.cuh which test_hostmath.py pins to the oracle bit for bit):

  * a `miss` verdict must never hit in the reference          (the conservative prefilter of every path),
  * a `certain` verdict must always hit in the reference      (the occupancy-only path skips the clip for those).

Regimes: voxel-sized triangles, leaf-sized ones, long thin leaves, slivers, vertices on voxel planes, axis-aligned
planes, grid-sized triangles, all at coordinates up to 8192 (the largest sample resolution)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT

SHIM = os.path.join(ROOT, "obj2voxel_b200", "libo2v_hostmath_test.so")
fp = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def shim():
    if not os.path.exists(SHIM):
        import obj2voxel_b200
        obj2voxel_b200.build()
    lib = C.CDLL(SHIM)
    lib.o2vt_classify_fuzz.argtypes = [fp, C.c_size_t, C.c_ulonglong, C.c_float, C.POINTER(C.c_ulonglong)]
    return lib


def certain_margin(grid):
    """certainMarginFor(S) of o2v_sat.cuh: what the kernels use at sample resolution S."""
    return 1.0 / 64 if grid <= 2048 else 1.0 / 32


def run(shim, leaves, grid=8192, max_volume=200_000):
    leaves = np.ascontiguousarray(leaves, np.float32).reshape(-1, 9)
    out = (C.c_ulonglong * 16)()
    shim.o2vt_classify_fuzz(leaves.ctypes.data_as(fp), len(leaves), max_volume, certain_margin(grid), out)
    keys = ["pairs", "miss", "uncertain", "certain", "hits", "miss_but_hit", "certain_but_no_hit", "skipped",
            "span_miss", "span_uncertain", "span_certain", "span_miss_but_hit", "span_certain_but_no_hit",
            "span_differs"]
    return dict(zip(keys, [int(x) for x in out]))


def blobs(rng, n, grid, size):
    centre = rng.uniform(size + 1, grid - size - 1, (n, 1, 3))
    return (centre + rng.uniform(-size, size, (n, 3, 3))).reshape(n, 9)


def needles(rng, n, grid, length, width):
    v0 = rng.uniform(length + 1, grid - length - 1, (n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    perp = rng.normal(size=(n, 3))
    perp -= (perp * d).sum(1, keepdims=True) * d
    perp /= np.linalg.norm(perp, axis=1, keepdims=True)
    ln = rng.uniform(0.2, 1.0, (n, 1)) * length
    wd = 10.0 ** rng.uniform(np.log10(width[0]), np.log10(width[1]), (n, 1))
    return np.concatenate([v0, v0 + ln * d, v0 + rng.uniform(0.2, 0.8, (n, 1)) * ln * d + wd * perp], axis=1)


def snapped(rng, n, grid, size):
    """Vertices on voxel planes / edges / corners, whole triangles inside axis planes."""
    t = blobs(rng, n, grid, size).reshape(n, 3, 3)
    mode = rng.integers(0, 4, n)
    for i in range(n):
        if mode[i] == 0:
            t[i, rng.integers(0, 3), rng.integers(0, 3)] = np.floor(t[i, 0, 0])
        elif mode[i] == 1:
            t[i, :, rng.integers(0, 3)] = np.floor(t[i, 0, 0]) + rng.choice([0.0, 0.5, 2.0 ** -17, -2.0 ** -17])
        elif mode[i] == 2:
            t[i, rng.integers(0, 3)] = np.floor(t[i, 0])
        else:
            t[i] = np.round(t[i] * 2) / 2
    return t.reshape(n, 9)


def grid_sized(rng, n, grid):
    """Huge, roughly axis-aligned triangles (the ones the reference does not subdivide): flat AABBs."""
    t = rng.uniform(1, grid - 1, (n, 3, 3))
    axis = rng.integers(0, 3, n)
    for i in range(n):
        base = rng.uniform(2, grid - 2)
        t[i, :, axis[i]] = base + rng.uniform(-0.4, 0.4, 3) * rng.choice([0.0, 1.0, 0.01])
        # keep the in-plane extent moderate so that the sweep stays affordable
        c = t[i].mean(axis=0, keepdims=True)
        t[i] = c + (t[i] - c) * min(1.0, 150.0 / (np.abs(t[i] - c).max() + 1e-9))
        t[i, :, axis[i]] = np.clip(t[i, :, axis[i]], 1, grid - 1)
    return t.reshape(n, 9)


REGIMES = [
    ("voxel-sized @64", 64, lambda r: blobs(r, 4000, 64, 1.2)),
    ("voxel-sized @2048", 2048, lambda r: blobs(r, 4000, 2048, 1.2)),
    ("voxel-sized @8192", 8192, lambda r: blobs(r, 4000, 8192, 1.2)),
    ("cfg4-sized @2048", 2048, lambda r: blobs(r, 3000, 2048, 2.05)),
    ("leaf-sized @2048", 2048, lambda r: blobs(r, 1500, 2048, 4.0)),
    ("leaf-sized @8192", 8192, lambda r: blobs(r, 1500, 8192, 4.0)),
    ("large @2048", 2048, lambda r: blobs(r, 60, 2048, 25.0)),
    ("large @8192", 8192, lambda r: blobs(r, 60, 8192, 25.0)),
    ("needles @2048", 2048, lambda r: needles(r, 600, 2048, 60.0, (1e-7, 1e-1))),
    ("needles @8192", 8192, lambda r: needles(r, 600, 8192, 200.0, (1e-6, 1.0))),
    ("snapped @256", 256, lambda r: snapped(r, 3000, 256, 2.5)),
    ("snapped @2048", 2048, lambda r: snapped(r, 3000, 2048, 2.5)),
    ("snapped @8192", 8192, lambda r: snapped(r, 3000, 8192, 2.5)),
    ("grid-sized @2048", 2048, lambda r: grid_sized(r, 60, 2048)),
    ("grid-sized @8192", 8192, lambda r: grid_sized(r, 60, 8192)),
]


@pytest.mark.parametrize("name,grid,make", REGIMES, ids=[r[0] for r in REGIMES])
def test_verdicts_agree_with_the_reference_semantics(shim, name, grid, make):
    """Both forms of the classifier — per voxel (thread-per-leaf kernel, weighted-path prefilter) and per row
    (classifySpan: the block classifier) — with the `certain` margin the kernels use at that grid size."""
    rng = np.random.default_rng(sum(map(ord, name)))
    stats = run(shim, make(rng), grid)
    assert stats["pairs"] > 1000, stats
    assert stats["miss_but_hit"] == 0, (name, stats)
    assert stats["certain_but_no_hit"] == 0, (name, stats)
    assert stats["span_miss_but_hit"] == 0, (name, stats)
    assert stats["span_certain_but_no_hit"] == 0, (name, stats)
    # the two forms differ only by rounding at a threshold: a handful of voxels at most
    assert stats["span_differs"] <= max(4, stats["pairs"] // 100_000), (name, stats)
    if "needles" not in name:
        assert stats["certain"] > 0 and stats["span_certain"] > 0, (name, stats)  # the sweep exercises the verdicts


def test_certain_covers_most_hits_on_the_bench_workload(shim):
    """The occupancy-only path pays an exact clip only for the `uncertain` band: on cfg4-like triangles it must be a small
    fraction of the hits (this is a performance property, pinned loosely)."""
    stats = run(shim, blobs(np.random.default_rng(3), 5000, 2048, 2.05), 2048)
    assert stats["certain"] > 0.8 * stats["hits"], stats
    assert stats["span_certain"] > 0.8 * stats["hits"], stats
