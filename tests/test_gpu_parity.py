"""GPU (-m gpu): parity of the CUDA path, called through the C-ABI, against the golden fixtures (generated from the
unmodified reference), against the CPU oracle on seeded inputs, and through size-independent properties at BASELINE
sizes.  Integer outputs (positions, ARGB8) must be bit-exact."""
import zlib

import numpy as np
import pytest

from conftest import golden_names, gpu_run, load_golden, oracle_kwargs
import obj2voxel_b200 as o2v
from obj2voxel_b200 import _lib, meshes
from oracle import oracle

pytestmark = pytest.mark.gpu

PATCHED = [n for n in golden_names() if "patched" in n]
EXACT = [n for n in golden_names() if "patched" not in n]


# ---- the reference's own tests (test/main.cpp), through the same public API ---------------------------------------

def run_instance(verts, resolution, **kw):
    inst = o2v.Instance()
    inst.set_input_callback(verts, uvs=kw.get("uvs"), texture=kw.get("texture"))
    inst.set_output_callback()
    inst.set_resolution(resolution)
    if "strategy" in kw:
        inst.set_color_strategy(kw["strategy"])
    if "supersampling" in kw:
        inst.set_supersampling(kw["supersampling"])
    if "bounds" in kw:
        inst.set_mesh_boundaries(kw["bounds"])
    err = inst.voxelize()
    voxels = inst.collected()
    inst.free()
    return err, voxels


def expected_unit_cube_voxels(r):
    return 8 + 12 * (r - 2) + 6 * (r - 2) * (r - 2)


def test_unit_cube_produces_expected_voxel_count():  # test/main.cpp:128-156
    err, voxels = run_instance(meshes.unit_cube(), 64)
    assert err == o2v.ERR_OK and len(voxels) == expected_unit_cube_voxels(64) == 23816


def test_unit_cube_produces_expected_byte_count():  # test/main.cpp:158-179
    inst = o2v.Instance()
    inst.set_input_callback(meshes.unit_cube())
    inst.set_output_memory("vl32")
    inst.set_resolution(64)
    assert inst.voxelize() == o2v.ERR_OK
    data = inst.get_output_memory()
    inst.free()
    assert len(data) == expected_unit_cube_voxels(64) * 16
    quads = np.frombuffer(data, dtype=">u4").reshape(-1, 4)  # VL32: big-endian x, y, z, argb
    assert quads[:, :3].max() == 63 and np.all(quads[:, 3] == 0xFFFFFFFF)


def test_unit_cube_multiple_chunks():  # test/main.cpp:194-208
    err, voxels = run_instance(meshes.unit_cube(), 128)
    assert err == o2v.ERR_OK and len(voxels) == expected_unit_cube_voxels(128) == 96776


@pytest.mark.parametrize("resolution", [32, 128])
def test_three_planes(resolution):  # test/main.cpp:225-252
    err, voxels = run_instance(meshes.three_planes(), resolution)
    assert err == o2v.ERR_OK and len(voxels) == 3 * resolution * resolution


def test_cfg1_single_triangle_golden_set():  # BASELINE config 1, SURVEY §8c
    err, voxels = run_instance(meshes.single_triangle(), 16)
    want = sorted((x, 0, z) for x in range(16) for z in range(16) if x + z <= 15)
    assert err == o2v.ERR_OK
    assert [tuple(v[:3]) for v in voxels.tolist()] == want and np.all(voxels[:, 3] == 0xFFFFFFFF)


def test_double_voxelization_and_sink_failure():  # src/obj2voxel.cpp:604-606,509-512
    o2v.load().obj2voxel_set_log_level(_lib.LOG_SILENT)
    inst = o2v.Instance()
    inst.set_input_callback(meshes.unit_cube())
    inst.set_output_callback()
    inst.set_resolution(16)
    assert inst.voxelize() == o2v.ERR_OK
    assert inst.voxelize() == o2v.ERR_DOUBLE_VOXELIZATION
    inst.free()
    inst = o2v.Instance()
    inst.set_input_callback(meshes.unit_cube())
    inst.set_output_callback(fail_after=0)  # the sink refuses the first batch
    inst.set_resolution(16)
    assert inst.voxelize() == o2v.ERR_IO_WRITE
    inst.free()
    o2v.load().obj2voxel_set_log_level(_lib.LOG_INFO)


def test_empty_model_is_ok_and_empty():  # src/obj2voxel.cpp:590-594
    o2v.load().obj2voxel_set_log_level(_lib.LOG_SILENT)
    err, voxels = run_instance(np.zeros((0, 9), np.float32), 16)
    o2v.load().obj2voxel_set_log_level(_lib.LOG_INFO)
    assert err == o2v.ERR_OK and len(voxels) == 0


@pytest.mark.parametrize("textured", [False, True])
@pytest.mark.parametrize("parts", [2, 4, 7])
def test_job_in_z_parts_equals_job_in_one_piece(monkeypatch, parts, textured):
    """obj2voxel_voxelize() runs big jobs as z sub-slabs so that the download of one part overlaps the kernels of the
    next (o2v_capi.cpp); forced here on a small mesh: the sink must receive exactly the same records (both pipelines),
    and the oracle's."""
    verts = meshes.random_triangles(4000, 0.03, seed=11)
    kw = dict(strategy=o2v.BLEND_STRATEGY, bounds=meshes.UNIT_BOUNDS)
    okw = dict(strategy=o2v.BLEND_STRATEGY, bounds=meshes.UNIT_BOUNDS)
    if textured:
        uvs = meshes.random_uvs(4000, seed=12)
        pixels = meshes.random_texture(32, 16, 3)
        kw.update(uvs=uvs, texture=o2v.Texture(pixels, wrap=o2v.UV_WRAP))
        okw.update(uvs=uvs, texture=dict(pixels=pixels, wrap=o2v.UV_WRAP))
    monkeypatch.setenv("O2V_B200_PIPELINE_PARTS", "1")
    err, whole = run_instance(verts, 200, **kw)
    assert err == o2v.ERR_OK
    monkeypatch.setenv("O2V_B200_PIPELINE_PARTS", str(parts))
    err, pieces = run_instance(verts, 200, **kw)
    assert err == o2v.ERR_OK
    whole, pieces = o2v.sort_voxels(whole), o2v.sort_voxels(pieces)
    assert np.array_equal(whole, pieces)
    assert np.array_equal(whole, oracle.voxelize(verts, 200, **okw)["voxels"])


def test_colored_triangles_voxelize_white_like_the_reference():
    inst = o2v.Instance()
    inst.set_input_callback(meshes.single_triangle(), colors=np.array([[1.0, 0.0, 0.0]], np.float32))
    inst.set_output_callback()
    inst.set_resolution(16)
    assert inst.voxelize() == o2v.ERR_OK
    assert np.all(inst.collected()[:, 3] == 0xFFFFFFFF)
    inst.free()


def test_reference_test_program_passes_against_this_library():
    """The reference's own test/main.cpp + testutil.hpp, compiled unmodified and linked against libobj2voxel_b200.so
    (oracle/Makefile target _ref/obj2voxel-test-b200; built where /root/reference exists, the binary travels)."""
    import os
    import subprocess

    from conftest import ROOT

    exe = os.path.join(ROOT, "oracle", "_ref", "obj2voxel-test-b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/obj2voxel-test-b200 not built")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "All tests passed" in out.stdout


# ---- golden fixtures from the unmodified reference -----------------------------------------------------------------

@pytest.mark.parametrize("name", EXACT)
def test_golden_bit_exact(engine, name):
    g = load_golden(name)
    got, stats = gpu_run(engine, g)
    if "api_voxels" in g:
        want = g["api_voxels"]
    else:  # UNTEXTURED colours cannot go through the reference's public API; the pinned oracle stands in
        want = oracle.voxelize(g["verts"], int(g["resolution"]), **oracle_kwargs(g))["voxels"]
    assert np.array_equal(np.asarray(stats["transform"], np.float32).view(np.uint32), g["transform_bits"])
    assert got.shape == want.shape and np.array_equal(got, want)
    assert np.array_equal(got[:, :3], g["int_xyz"])
    materialless = "uvs" not in g and "types" not in g
    assert bool(stats["occupancy_path"]) == materialless
    if materialless:
        # all-white meshes take the occupancy-only path by default; the weighted path and the occupancy path without
        # its SAT shortcuts (every candidate through the exact clip) must give the very same records
        for kw in (dict(occupancy_path=0), dict(prefilter=0)):
            other, other_stats = gpu_run(engine, g, **kw)
            assert bool(other_stats["occupancy_path"]) == ("occupancy_path" not in kw)
            assert np.array_equal(other, want)


@pytest.mark.parametrize("name", PATCHED)
def test_golden_supersampling(engine, name):
    """Occupancy exact; colours exact for MAX and within 1 LSB for BLEND (child fold order of the patched reference is
    unordered_map order; ours is ascending Morton)."""
    g = load_golden(name)
    got, stats = gpu_run(engine, g, occupancy_path=0)
    fast, fast_stats = gpu_run(engine, g)
    assert not stats["occupancy_path"] and fast_stats["occupancy_path"]
    assert np.array_equal(fast, got)  # the occupancy-only path (default for this all-white mesh) changes nothing
    want = g["api_voxels"]
    assert np.array_equal(got[:, :3], want[:, :3])
    tol = 0 if int(g["strategy"]) == 0 else 1
    for shift in (0, 8, 16, 24):
        a = (got[:, 3].astype(np.int64) >> shift) & 255
        b = (want[:, 3].astype(np.int64) >> shift) & 255
        assert np.max(np.abs(a - b)) <= tol
    # and bit-exact against the oracle, which defines the fold order
    ref = oracle.voxelize(g["verts"], int(g["resolution"]), **oracle_kwargs(g))["voxels"]
    assert np.array_equal(got, ref)


# ---- seeded parity against the oracle ------------------------------------------------------------------------------

def parity(engine, verts, resolution, uvs=None, texture=None, types=None, colors=None, **kw):
    textures = [(texture["pixels"], texture["wrap"])] if texture is not None else []
    okw = {k: v for k, v in kw.items() if k in ("strategy", "supersampling", "bounds", "unit")}
    want = oracle.voxelize(verts, resolution, uvs=uvs, texture=texture, types=types, colors=colors, **okw)
    # weighted path (the only one for coloured / textured meshes)
    params = o2v.make_params(resolution=resolution, occupancy_path=0, **kw)
    got, stats = engine.voxelize_host(verts, params, uvs=uvs, types=types, colors=colors, textures=textures)
    got = o2v.sort_voxels(got)
    assert got.shape == want["voxels"].shape and np.array_equal(got, want["voxels"])
    assert stats["contributions"] == want["contributions"]  # N_contrib agrees with the reference's emplace count
    if texture is None and types is None:
        # all-white mesh: the default is the occupancy-only path
        fast, fast_stats = engine.voxelize_host(verts, o2v.make_params(resolution=resolution, **kw))
        assert fast_stats["occupancy_path"] and not stats["occupancy_path"]
        assert np.array_equal(o2v.sort_voxels(fast), want["voxels"])
        assert fast_stats["clip_calls"] <= stats["clip_calls"]
    return stats


def test_cfg2_lumpy_sphere_r256_max(engine):  # BASELINE config 2 (70 k triangles)
    parity(engine, meshes.lumpy_sphere(), 256, strategy=0)


@pytest.mark.parametrize("strategy", [0, 1])
def test_random_textured_r256(engine, strategy):  # BASELINE config 3 scaled to oracle-friendly size
    n = 60000
    v = meshes.random_triangles(n, 0.008, seed=1)
    uv = meshes.random_uvs(n, seed=2)
    tex = dict(pixels=meshes.random_texture(256, 256, 3), wrap=o2v.UV_WRAP)
    parity(engine, v, 256, uvs=uv, texture=tex, strategy=strategy, bounds=[-0.01, -0.01, -0.01, 1.01, 1.01, 1.01])


def test_micro_triangles_r512(engine):  # BASELINE config 5 regime: sub-voxel triangles, atomic-free fold
    v = meshes.random_triangles(200000, 0.25 / 512, seed=7)
    parity(engine, v, 512, strategy=0, bounds=[-0.01, -0.01, -0.01, 1.01, 1.01, 1.01])


def test_large_triangles_deep_subdivision(engine):
    parity(engine, meshes.random_triangles(40, 0.45, seed=17), 512, strategy=1)


def test_huge_triangles_are_walked_in_reference_order(engine):
    """Triangles whose voxel AABB passes 2^21 voxels are only listed by the count pass and subdivided as 256 (triangle,
    subtree) items spread over the device (o2v_device.cuh, forEachHugeLeaf): textured + BLEND makes every voxel depend on
    the order of the leaves, the small triangles around them take the ordinary walk, supersampling adds the downscale."""
    big = meshes.random_triangles(6, 0.45, seed=29)
    small = meshes.random_triangles(3000, 0.01, seed=30)
    rng = np.random.default_rng(31)
    order = rng.permutation(len(big) + len(small))
    v = np.concatenate([big, small])[order]
    uv = meshes.random_uvs(len(v), seed=32)
    tex = dict(pixels=meshes.random_texture(64, 64, 3, seed=33), wrap=o2v.UV_WRAP)
    box = [-0.01, -0.01, -0.01, 1.01, 1.01, 1.01]
    stats = parity(engine, v, 384, uvs=uv, texture=tex, strategy=1, bounds=box)
    assert stats["leaves"] > len(v) + 1000, stats["leaves"]  # the big ones really subdivide
    parity(engine, v, 160, uvs=uv, texture=tex, strategy=0, supersampling=2, bounds=box)
    parity(engine, v, 384, strategy=0, bounds=box)  # all-white: both pipelines


@pytest.mark.parametrize("supersampling", [1, 2])
def test_pieces_of_a_mesh_accumulate_to_the_whole(engine, supersampling):
    """o2v_b200_params::accumulate: the mesh in three pieces on one engine — every piece returns the voxels no earlier piece
    has produced, the pieces' results are disjoint, their union is the whole mesh's result."""
    v = meshes.random_triangles(30000, 0.02, seed=71)
    res = 256 // supersampling
    kw = dict(resolution=res, supersampling=supersampling, strategy=0, bounds=meshes.UNIT_BOUNDS)
    want = oracle.voxelize(v, res, supersampling=supersampling, strategy=0, bounds=meshes.UNIT_BOUNDS)["voxels"]
    got = []
    for k, piece in enumerate(np.array_split(v, 3)):
        # (room for every record at once: a second call for the same piece would find all of its voxels delivered)
        records, stats = engine.voxelize_host(np.ascontiguousarray(piece), o2v.make_params(accumulate=1 if k == 0 else 2, **kw),
                                              capacity=1 << 21)
        assert stats["occupancy_path"] == 1
        got.append(records)
    assert sum(len(g) for g in got) == len(want) and len(got[1]) < len(got[0])
    assert np.array_equal(o2v.sort_voxels(np.concatenate(got)), want)
    whole, _ = engine.voxelize_host(v, o2v.make_params(**kw))  # an ordinary run afterwards
    assert np.array_equal(o2v.sort_voxels(whole), want)


def test_huge_triangle_list_grows_and_shrinks_between_runs(engine):
    """The list of huge triangles is sized by what the previous attempt met: runs with none, a few, many and none again on
    one engine (every change of the count is a run that starts over once)."""
    box = [-0.01, -0.01, -0.01, 1.01, 1.01, 1.01]
    for n_big, seed in ((0, 51), (3, 52), (40, 53), (0, 54), (5, 55)):
        parts = [meshes.random_triangles(300, 0.02, seed=seed)]
        if n_big:
            parts.append(meshes.random_triangles(n_big, 0.45, seed=seed + 100))
        v = np.concatenate(parts)
        want = oracle.voxelize(v, 256, strategy=0, bounds=box)["voxels"]
        for occupancy in (1, 0):
            got, _ = engine.voxelize_host(v, o2v.make_params(resolution=256, strategy=0, bounds=box, occupancy_path=occupancy))
            assert np.array_equal(o2v.sort_voxels(got), want), (n_big, occupancy)


def test_huge_triangle_in_a_slab(engine):
    v = np.concatenate([meshes.random_triangles(3, 0.45, seed=41), meshes.random_triangles(500, 0.02, seed=42)])
    box = [-0.01, -0.01, -0.01, 1.01, 1.01, 1.01]
    want = oracle.voxelize(v, 256, strategy=0, bounds=box)["voxels"]
    parts = []
    for slab in ((0, 64), (64, 192), (192, 256)):
        for occupancy in (0, 1):
            got, _ = engine.voxelize_host(v, o2v.make_params(resolution=256, strategy=0, bounds=box, slab=slab,
                                                             occupancy_path=occupancy))
            got = o2v.sort_voxels(got)
            expect = want[(want[:, 2] >= slab[0]) & (want[:, 2] < slab[1])]
            assert np.array_equal(got, expect), (slab, occupancy)


def test_many_triangles_in_one_tile_long_lists(engine):
    """20 k triangles crowded into a few tiles: exercises the large-list sort and > 32-leaf batches."""
    v = meshes.random_triangles(20000, 0.02, seed=23) * np.float32(0.1)
    parity(engine, v, 32, strategy=1, bounds=[0, 0, 0, 1, 1, 1])


def test_mixed_materials(engine):
    rng = np.random.default_rng(4)
    n = 5000
    v = meshes.random_triangles(n, 0.03, seed=9)
    types = rng.integers(1, 4, n).astype(np.uint8)
    cols = rng.random((n, 3)).astype(np.float32)
    uv = meshes.random_uvs(n, seed=10) * 2 - 0.5
    tex = dict(pixels=meshes.random_texture(32, 32, 4, seed=6), wrap=o2v.UV_CLAMP)
    parity(engine, v, 128, uvs=uv, texture=tex, types=types, colors=cols, strategy=1,
           bounds=[-0.05, -0.05, -0.05, 1.05, 1.05, 1.05])


def test_out_of_bounds_triangles_match_reference_behaviour(engine):
    """Mesh bounds smaller than the mesh: negative voxel coordinates drop the triangle, coordinates beyond the resolution
    survive up to the 64^3 chunk grid — the observed behaviour of the reference, pinned in the oracle."""
    v = meshes.random_triangles(800, 0.08, seed=21)
    for res, b in ((16, [0.2, 0.2, 0.2, 0.7, 0.7, 0.7]), (100, [0.1, 0.1, 0.1, 0.95, 0.9, 0.9])):
        parity(engine, v, res, strategy=1, bounds=b)


def test_supersampling_pre_downscale_equals_double_resolution(engine):
    g = load_golden("rand500_r64_presample_of_ss2")
    got, _ = gpu_run(engine, g)  # resolution 64, ss 1 == the sample grid of (32, ss 2)
    assert np.array_equal(got, g["api_voxels"])


# ---- occupancy-only path: the pieces the fixtures above do not reach --------------------------------------------------

def test_unit_cube_r256_golden_count_and_big_leaf_boxes(engine):
    """SURVEY §8c golden: unit cube at 256 -> 390,152 voxels.  Its 12 axis-aligned triangles are not subdivided: 65,536+
    candidates each, classified box by box (16^3) on the occupancy-only path and by the heavy-tile kernel otherwise."""
    for occ in (1, 0):
        got, stats = engine.voxelize_host(meshes.unit_cube(), o2v.make_params(resolution=256, occupancy_path=occ))
        assert bool(stats["occupancy_path"]) == bool(occ)
        assert len(got) == 390152 and np.all(got[:, 3] == 0xFFFFFFFF)
        s = o2v.sort_voxels(got)
        on_face = (s[:, :3] == 0).any(axis=1) | (s[:, :3] == 255).any(axis=1)
        assert on_face.all()


def test_occupancy_slab_union_equals_whole(engine):
    """Z-slabs that are multiples of 8 but not of the 64-voxel bitmap chunks; supersampled and not."""
    v = meshes.random_triangles(150000, 0.006, seed=77)
    b = [-0.01, -0.01, -0.01, 1.01, 1.01, 1.01]
    for res, ss, cuts in ((256, 1, (0, 72, 200, 256)), (128, 2, (0, 8, 136, 256))):
        whole, stats = engine.voxelize_host(v, o2v.make_params(resolution=res, supersampling=ss, bounds=b))
        assert stats["occupancy_path"]
        weighted, _ = engine.voxelize_host(v, o2v.make_params(resolution=res, supersampling=ss, bounds=b,
                                                              occupancy_path=0))
        assert checksum(whole) == checksum(weighted)
        parts = []
        for z0, z1 in zip(cuts[:-1], cuts[1:]):
            part, _ = engine.voxelize_host(v, o2v.make_params(resolution=res, supersampling=ss, bounds=b,
                                                              slab=(z0, z1)))
            assert len(part) == 0 or (part[:, 2].min() >= z0 // ss and part[:, 2].max() < -(-z1 // ss))
            parts.append(part)
        assert checksum(np.concatenate(parts)) == checksum(whole)


def test_occupancy_queue_overflow_reruns_with_a_larger_queue(engine):
    """prefilter = 0 queues every candidate voxel: more than the initial queue (a quarter of the candidates), so the
    engine must grow it and rerun — same records."""
    v = meshes.random_triangles(80000, 0.01, seed=5)
    b = [-0.02, -0.02, -0.02, 1.02, 1.02, 1.02]
    fast, fs = engine.voxelize_host(v, o2v.make_params(resolution=256, bounds=b))
    slow, ss = engine.voxelize_host(v, o2v.make_params(resolution=256, bounds=b, prefilter=0))
    assert fs["occupancy_path"] and ss["occupancy_path"]
    assert ss["clip_calls"] > 4 * fs["clip_calls"] and ss["candidate_voxels"] > (1 << 21)
    assert checksum(fast) == checksum(slow)


def test_occupancy_falls_back_when_the_bitmaps_do_not_fit(engine, monkeypatch):
    v = meshes.random_triangles(20000, 0.01, seed=6)
    b = [-0.02, -0.02, -0.02, 1.02, 1.02, 1.02]
    want, ws = engine.voxelize_host(v, o2v.make_params(resolution=128, bounds=b))
    monkeypatch.setenv("O2V_B200_OCCUPANCY_MAX_BYTES", "1000")
    got, gs = engine.voxelize_host(v, o2v.make_params(resolution=128, bounds=b))
    assert ws["occupancy_path"] and not gs["occupancy_path"]
    assert checksum(got) == checksum(want)


# ---- properties at BASELINE sizes (no oracle needed) ---------------------------------------------------------------

def checksum(voxels):
    return zlib.crc32(np.ascontiguousarray(o2v.sort_voxels(voxels)).tobytes())


@pytest.fixture(scope="module")
def cfg3_run(engine):
    cfg = meshes.CONFIGS["cfg3"]
    n = cfg["n"]
    v = meshes.random_triangles(n, cfg["extent"], seed=1)
    uv = meshes.random_uvs(n, seed=2)
    tex = (meshes.random_texture(256, 256, 3), o2v.UV_WRAP)
    bounds = [-0.01, -0.01, -0.01, 1.01, 1.01, 1.01]
    params = o2v.make_params(resolution=cfg["resolution"], strategy=cfg["strategy"], bounds=bounds)
    voxels, stats = engine.voxelize_host(v, params, uvs=uv, textures=[tex], capacity=1 << 25)
    return dict(v=v, uv=uv, tex=tex, bounds=bounds, cfg=cfg, voxels=voxels, stats=stats)


def test_cfg3_full_size_is_deterministic_and_well_formed(engine, cfg3_run):
    r = cfg3_run
    params = o2v.make_params(resolution=r["cfg"]["resolution"], strategy=r["cfg"]["strategy"], bounds=r["bounds"])
    again, _ = engine.voxelize_host(r["v"], params, uvs=r["uv"], textures=[r["tex"]], capacity=1 << 25)
    assert checksum(again) == checksum(r["voxels"])  # ordered fold: run-to-run identical, unlike atomics
    s = o2v.sort_voxels(r["voxels"])
    assert s[:, :3].max() < r["cfg"]["resolution"]
    keys = (s[:, 0].astype(np.int64) << 40) | (s[:, 1].astype(np.int64) << 20) | s[:, 2].astype(np.int64)
    assert np.all(np.diff(keys) > 0)  # every voxel exactly once
    assert np.all((s[:, 3] >> 24) == 0xFF)


def test_cfg3_prefilter_is_conservative(engine, cfg3_run):
    r = cfg3_run
    params = o2v.make_params(resolution=r["cfg"]["resolution"], strategy=r["cfg"]["strategy"], bounds=r["bounds"],
                             prefilter=0)
    full, stats = engine.voxelize_host(r["v"], params, uvs=r["uv"], textures=[r["tex"]], capacity=1 << 25)
    assert checksum(full) == checksum(r["voxels"])
    assert stats["clip_calls"] > r["stats"]["clip_calls"]  # the filter did remove work, never results


def test_cfg3_slab_union_equals_whole(engine, cfg3_run):
    r = cfg3_run
    parts = []
    for z0, z1 in ((0, 128), (128, 320), (320, 512)):
        params = o2v.make_params(resolution=r["cfg"]["resolution"], strategy=r["cfg"]["strategy"], bounds=r["bounds"],
                                 slab=(z0, z1))
        part, _ = engine.voxelize_host(r["v"], params, uvs=r["uv"], textures=[r["tex"]], capacity=1 << 25)
        assert len(part) == 0 or (part[:, 2].min() >= z0 and part[:, 2].max() < z1)
        parts.append(part)
    assert checksum(np.concatenate(parts)) == checksum(r["voxels"])


def test_device_resident_path_equals_host_path(engine):
    import torch

    v = meshes.random_triangles(50000, 0.01, seed=3)
    params = o2v.make_params(resolution=256, strategy=1, bounds=[-0.02, -0.02, -0.02, 1.02, 1.02, 1.02])
    host, _ = engine.voxelize_host(v, params)
    dv = torch.from_numpy(v).cuda()
    same = meshes.random_triangles_torch(50000, 0.01, seed=3)
    assert torch.equal(dv, same)  # host and device generators are bit-identical
    engine.voxelize_device(dv, params)
    dev = engine.result_tensor().cpu().numpy().view(np.uint32)
    assert checksum(dev) == checksum(host)
    assert np.array_equal(o2v.sort_voxels(engine.download()), o2v.sort_voxels(host))


def test_triangle_array_that_is_not_16_byte_aligned(engine):
    """The count / emit / slab-filter passes stream triangle batches with bulk-async copies, which need 16-byte aligned
    source addresses; an array that is only 4-byte aligned (a view one float into a buffer) takes the block-load path and
    must give the same records — whole grid and slab (the slab filter reads it too)."""
    import torch

    v = meshes.random_triangles(30001, 0.01, seed=5)  # ragged last batch as well
    params = o2v.make_params(resolution=256, bounds=[-0.02, -0.02, -0.02, 1.02, 1.02, 1.02])
    want = o2v.sort_voxels(engine.voxelize_host(v, params)[0])
    buf = torch.zeros(v.size + 1, dtype=torch.float32, device="cuda")
    buf[1:] = torch.from_numpy(v).cuda().reshape(-1)
    shifted = buf[1:].view(-1, 9)
    assert shifted.data_ptr() % 16 == 4
    engine.voxelize_device(shifted, params)
    assert np.array_equal(o2v.sort_voxels(engine.download()), want)
    parts = []
    for slab in ((0, 64), (64, 256)):
        engine.voxelize_device(shifted, o2v.make_params(resolution=256, bounds=[-0.02, -0.02, -0.02, 1.02, 1.02, 1.02],
                                                        slab=slab))
        parts.append(engine.download())
    assert np.array_equal(o2v.sort_voxels(np.concatenate(parts)), want)


def test_more_chunks_than_a_block_collects(engine):
    """Above 65536 chunks per slab (grids beyond 2560^3) the count pass marks chunks in global memory instead of per
    block; 2688^3 = 42^3 = 74088 chunks, a few thousand small triangles: occupancy path vs weighted path vs oracle."""
    v = meshes.random_triangles(3000, 0.002, seed=9)
    kw = dict(resolution=2688, bounds=meshes.UNIT_BOUNDS)
    got, stats = engine.voxelize_host(v, o2v.make_params(**kw))
    assert stats["occupancy_path"]
    weighted, wstats = engine.voxelize_host(v, o2v.make_params(occupancy_path=0, **kw))
    assert not wstats["occupancy_path"]
    got = o2v.sort_voxels(got)
    assert np.array_equal(got, o2v.sort_voxels(weighted))
    assert np.array_equal(got, oracle.voxelize(v, 2688, bounds=meshes.UNIT_BOUNDS)["voxels"])


@pytest.mark.parametrize("extent,resolution,supersampling", [(0.0005, 512, 1), (0.02, 128, 1), (0.004, 128, 2)])
def test_both_classifiers_agree_with_the_oracle(engine, extent, resolution, supersampling):
    """The occupancy pipeline classifies block = 64 leaves (variant 1) or, for meshes of micro-triangles, thread = leaf
    (variant 2; the engine picks by the average number of candidate voxels per leaf).  Forced both ways on micro and on
    ordinary triangles (big axis-aligned leaves ride along): identical records, the oracle's."""
    v = np.concatenate([meshes.random_triangles(20000, extent, seed=31), meshes.unit_cube() * np.float32(0.9) + np.float32(0.05)])
    kw = dict(resolution=resolution, supersampling=supersampling, bounds=meshes.UNIT_BOUNDS)
    want = oracle.voxelize(v, resolution, supersampling=supersampling, bounds=meshes.UNIT_BOUNDS)["voxels"]
    for variant in (1, 2):
        got, stats = engine.voxelize_host(v, o2v.make_params(variant=variant, **kw))
        assert stats["occupancy_path"]
        assert np.array_equal(o2v.sort_voxels(got), want), variant
