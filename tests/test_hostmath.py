"""CPU: the exact-arithmetic header the kernels run (obj2voxel_b200/csrc/o2v_exact.cuh), compiled for the host by the
test shim libo2v_hostmath_test.so, against the oracle — bit for bit.  This is how kernel arithmetic is checked on a
machine without a GPU; the shim is test infrastructure and is never loaded by the product."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle

SHIM = os.path.join(ROOT, "obj2voxel_b200", "libo2v_hostmath_test.so")
fp = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def shim():
    if not os.path.exists(SHIM):
        import obj2voxel_b200
        obj2voxel_b200.build()
    lib = C.CDLL(SHIM)
    lib.o2vt_clip_voxel.restype = C.c_int
    lib.o2vt_clip_voxel.argtypes = [fp, C.POINTER(C.c_uint32), C.c_float, C.c_int, fp]
    lib.o2vt_subdivide.restype = C.c_size_t
    lib.o2vt_subdivide.argtypes = [fp, fp, C.c_size_t]
    lib.o2vt_subdivide_by_subtrees.restype = C.c_size_t
    lib.o2vt_subdivide_by_subtrees.argtypes = [fp, C.c_int, fp, C.c_size_t]
    lib.o2vt_area.restype = C.c_float
    lib.o2vt_area.argtypes = [fp]
    lib.o2vt_transform.argtypes = [fp, fp, fp]
    lib.o2vt_texture_lookup.argtypes = [C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                        C.c_float, C.c_float, fp]
    lib.o2vt_quantize.restype = C.c_uint32
    lib.o2vt_quantize.argtypes = [fp]
    lib.o2vt_combine.argtypes = [fp, fp, C.c_int]
    return lib


def random_leaf(rng, it):
    c = rng.random(3) * 20 + 2
    scale = rng.choice([0.3, 1.0, 3.0, 8.0])
    tri = np.zeros(15, np.float32)
    tri[:9] = (c[None, :] + (rng.random((3, 3)) * 2 - 1) * scale).reshape(9)
    if it % 7 == 0:
        tri[0] = np.floor(tri[0])  # a vertex exactly on a voxel plane (planar cases of splitTriangle)
    if it % 11 == 0:
        tri[3] = tri[0]
    if it % 13 == 0:
        tri[[2, 5, 8]] = np.floor(tri[2])  # whole triangle inside a z plane
    tri[9:] = rng.random(6) * 4 - 2
    pos = np.floor(c + (rng.random(3) * 2 - 1) * scale * 0.7).astype(np.uint32)
    return tri, pos


def test_clip_matches_oracle_bitwise(shim):
    rng = np.random.default_rng(5)
    nonzero = 0
    for it in range(6000):
        tri, pos = random_leaf(rng, it)
        area = np.float32(rng.random() * 5 + 0.1)
        n_ref, w_ref = oracle.clip_voxel(tri, pos, area)
        out = np.zeros(3, np.float32)
        p = (C.c_uint32 * 3)(*pos.tolist())
        for textured in (1, 0):
            n = shim.o2vt_clip_voxel(tri.ctypes.data_as(fp), p, float(area), textured, out.ctypes.data_as(fp))
            assert n == n_ref
            assert out[0].view(np.uint32) == w_ref[0].view(np.uint32)
            if textured and n_ref > 0:
                assert np.array_equal(out.view(np.uint32), w_ref.view(np.uint32))
        nonzero += n_ref > 0
    assert nonzero > 1000  # the sweep really exercises surviving pieces


def test_subdivision_order_and_values(shim):
    rng = np.random.default_rng(6)
    for it in range(120):
        c = rng.random(3) * 200 + 20
        scale = rng.choice([1.0, 5, 30, 100])
        tri = np.zeros(15, np.float32)
        tri[:9] = (c[None, :] + (rng.random((3, 3)) * 2 - 1) * scale).reshape(9)
        tri[9:] = rng.random(6)
        want = oracle.subdivide(tri)
        got = np.zeros((max(len(want), 1), 15), np.float32)
        n = shim.o2vt_subdivide(tri.ctypes.data_as(fp), got.ctypes.data_as(fp), len(got))
        assert n == len(want)
        assert np.array_equal(got[:n].view(np.uint32), want.view(np.uint32))


def test_subtree_walk_visits_the_leaves_in_the_reference_order(shim):
    """Huge triangles are subdivided as 4^depth independent subtrees (o2v_exact.cuh, forEachLeafOfSubtree): walking the
    subtrees one after the other must give the oracle's leaves in the oracle's order — including triangles so small that
    they (or their children) are leaves above the split depth."""
    rng = np.random.default_rng(12)
    checked = 0
    for it in range(60):
        scale = [0.5, 3.0, 12.0, 40.0, 90.0][it % 5]
        tri = np.zeros(15, np.float32)
        tri[:9] = (rng.random(3)[None, :] * 50 + 60 + (rng.random((3, 3)) * 2 - 1) * scale).reshape(9)
        tri[9:] = rng.random(6)
        if shim.o2vt_aligned(tri.ctypes.data_as(fp)):
            continue  # (an aligned triangle is its own leaf and never takes this path)
        want = oracle.subdivide(tri)
        for depth in (1, 3, 4):
            got = np.zeros((max(len(want), 1), 15), np.float32)
            n = shim.o2vt_subdivide_by_subtrees(tri.ctypes.data_as(fp), depth, got.ctypes.data_as(fp), len(got))
            assert n == len(want), (it, depth, n, len(want))
            assert np.array_equal(got[:n].view(np.uint32), want.view(np.uint32)), (it, depth)
        checked += 1
    assert checked > 40


def test_transform_and_area(shim):
    rng = np.random.default_rng(8)
    m = oracle.mesh_transform([-1.5, 0.25, 3.0], [2.5, 1.75, 9.0], 1024, [0, 1, 0, 0, 0, -1, 1, 0, 0])
    for _ in range(200):
        v = (rng.random(3) * 6 - 1.5).astype(np.float32)
        out = np.zeros(3, np.float32)
        shim.o2vt_transform(m.ctypes.data_as(fp), v.ctypes.data_as(fp), out.ctypes.data_as(fp))
        # oracle: same affine through a degenerate one-triangle run is overkill; restate the row-wise dot in numpy f32
        want = np.zeros(3, np.float32)
        for i in range(3):
            r = np.float32(0)
            for j in range(3):
                r = np.float32(r + np.float32(m[i * 3 + j] * v[j]))
            want[i] = np.float32(r + m[9 + i])
        assert np.array_equal(out.view(np.uint32), want.view(np.uint32))


def test_texture_lookup_quantize_combine(shim):
    rng = np.random.default_rng(9)
    for channels, wrap in ((3, 1), (4, 0), (4, 1), (3, 0)):
        pixels = rng.integers(0, 256, (13, 21, channels), dtype=np.uint8)
        tex, _keep = oracle.make_texture(dict(pixels=pixels, wrap=wrap))
        for _ in range(300):
            uv = (rng.random(2) * 6 - 3).astype(np.float32)
            if rng.random() < 0.1:
                uv = np.round(uv)  # exact integers hit the REPEAT quirk
            want = np.zeros(3, np.float32)
            oracle.lib().o2v_oracle_texture_lookup(C.byref(tex), uv.ctypes.data_as(fp), want.ctypes.data_as(fp))
            got = np.zeros(3, np.float32)
            shim.o2vt_texture_lookup(pixels.ctypes.data_as(C.POINTER(C.c_uint8)), 21, 13, channels, wrap,
                                     float(uv[0]), float(uv[1]), got.ctypes.data_as(fp))
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    for _ in range(500):
        rgb = (rng.random(3) * 1.4 - 0.2).astype(np.float32)
        assert shim.o2vt_quantize(rgb.ctypes.data_as(fp)) == oracle.quantize_argb(rgb)
    # MAX keeps the existing value on ties; BLEND is the weighted mean in the reference's operand order
    acc = np.array([2.0, 0.1, 0.2, 0.3], np.float32)
    shim.o2vt_combine(acc.ctypes.data_as(fp), np.array([2.0, 0.9, 0.9, 0.9], np.float32).ctypes.data_as(fp), 0)
    assert acc.tolist() == pytest.approx([2.0, 0.1, 0.2, 0.3])
    shim.o2vt_combine(acc.ctypes.data_as(fp), np.array([6.0, 0.5, 0.6, 0.7], np.float32).ctypes.data_as(fp), 1)
    assert acc[0] == 8.0 and acc[1] == np.float32((np.float32(6.0) * np.float32(0.5) + np.float32(2.0) * np.float32(0.1)) / np.float32(8.0))


def test_magic_division_of_the_classify_kernel_is_exact():
    """o2v_occupancy.cu decodes a row-segment number with n / d == __umulhi(n, 0xffffffff / d + 1) (magicOf /
    divideBy); d <= 4096 (kOccBigVolume) and n * d <= 2^24 there: exhaustive over that range."""
    for d in range(2, 4097):
        magic = (0xFFFFFFFF // d + 1) & 0xFFFFFFFF
        n = np.arange(0, (1 << 24) // d + 1, dtype=np.uint64)
        assert np.array_equal((n * np.uint64(magic)) >> np.uint64(32), n // np.uint64(d)), d
