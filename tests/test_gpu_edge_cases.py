"""GPU (-m gpu): edge cases of the hot path — tiny grids, non-multiple-of-8 resolutions, degenerate and non-finite input,
single-point meshes, axis permutations, repeated runs on one engine with changing sizes."""
import numpy as np
import pytest

import obj2voxel_b200 as o2v
from obj2voxel_b200 import meshes
from oracle import oracle

pytestmark = pytest.mark.gpu


def both(engine, verts, resolution, **kw):
    """These meshes are all-white: the default run takes the occupancy-only path; the weighted path must agree."""
    got, stats = engine.voxelize_host(verts, o2v.make_params(resolution=resolution, **kw))
    weighted, wstats = engine.voxelize_host(verts, o2v.make_params(resolution=resolution, occupancy_path=0, **kw))
    assert stats["occupancy_path"] and not wstats["occupancy_path"]
    assert np.array_equal(o2v.sort_voxels(weighted), o2v.sort_voxels(got))
    want = oracle.voxelize(verts, resolution, **kw)["voxels"]
    return o2v.sort_voxels(got), want, stats


@pytest.mark.parametrize("resolution", [1, 2, 3, 7, 9, 17, 63, 65, 100])
def test_small_and_odd_resolutions(engine, resolution):
    for mesh in (meshes.unit_cube(), meshes.lumpy_sphere(12, 13)):
        got, want, _ = both(engine, mesh, resolution, strategy=1)
        assert np.array_equal(got, want)


def test_degenerate_and_non_finite_triangles_are_ignored(engine):
    good = meshes.random_triangles(300, 0.05, seed=5)
    bad = np.array([[0.5, 0.5, 0.5] * 3,                                  # point
                    [0.1, 0.1, 0.1, 0.2, 0.2, 0.2, 0.3, 0.3, 0.3],        # collinear
                    [np.nan, 0, 0, 0, 1, 0, 0, 0, 1],
                    [np.inf, 0, 0, 0, 1, 0, 0, 0, 1]], dtype=np.float32)
    mixed = np.concatenate([good[:150], bad[:2], good[150:]])              # zero-area only: the oracle defines the result
    got, want, _ = both(engine, mixed, 64, strategy=1, bounds=[-0.1, -0.1, -0.1, 1.1, 1.1, 1.1])
    assert np.array_equal(got, want)
    # NaN / inf vertices are a contract violation in the reference; here they are dropped and must not disturb the rest
    nasty = np.concatenate([good, bad])
    out, stats = engine.voxelize_host(nasty, o2v.make_params(resolution=64, strategy=1,
                                                             bounds=[-0.1, -0.1, -0.1, 1.1, 1.1, 1.1]))
    clean = oracle.voxelize(good, 64, strategy=1, bounds=[-0.1, -0.1, -0.1, 1.1, 1.1, 1.1])["voxels"]
    assert np.array_equal(o2v.sort_voxels(out), clean)
    assert stats["dropped_triangles"] == 4


def test_single_point_mesh_does_not_crash(engine):
    tri = np.full((3, 9), 0.25, dtype=np.float32)
    out, _ = engine.voxelize_host(tri, o2v.make_params(resolution=16))
    assert len(out) == 0


@pytest.mark.parametrize("unit", [[0, 1, 0, 0, 0, 1, 1, 0, 0], [-1, 0, 0, 0, 1, 0, 0, 0, -1], [0, 0, 1, 0, -1, 0, 1, 0, 0]])
def test_unit_transforms(engine, unit):  # obj2voxel_set_unit_transform, reference CLI -p permutations
    got, want, _ = both(engine, meshes.lumpy_sphere(14, 15) * np.float32([1, 0.5, 0.25] * 3), 48, unit=unit, strategy=0)
    assert np.array_equal(got, want)


def test_engine_reuse_with_changing_sizes(engine):
    """Grow-only buffers: a large job, a tiny one, a large one with another layout — results stay exact."""
    for n, res, ss in ((40000, 256, 1), (5, 16, 1), (20000, 64, 2), (1, 512, 1)):
        v = meshes.random_triangles(n, 0.01, seed=n)
        got, want, _ = both(engine, v, res, supersampling=ss, strategy=1, bounds=[-0.02, -0.02, -0.02, 1.02, 1.02, 1.02])
        assert np.array_equal(got, want)


def test_texture_table_with_two_textures(engine):
    rng = np.random.default_rng(3)
    n = 2000
    v = meshes.random_triangles(n, 0.03, seed=41)
    uv = meshes.random_uvs(n, seed=42) * 2 - 0.5
    ids = rng.integers(0, 2, n).astype(np.uint32)
    tex = [(meshes.random_texture(16, 8, 3, seed=1), o2v.UV_WRAP), (meshes.random_texture(5, 7, 3, seed=2), o2v.UV_WRAP)]
    params = o2v.make_params(resolution=64, strategy=1, bounds=[-0.05, -0.05, -0.05, 1.05, 1.05, 1.05])
    got, _ = engine.voxelize_host(v, params, uvs=uv, texture_ids=ids, textures=tex)
    # the oracle takes one texture: voxelize each subset and check that every voxel hit by only one subset agrees
    parts = []
    for t in (0, 1):
        sel = ids == t
        r = oracle.voxelize(v[sel], 64, uvs=uv[sel], texture=dict(pixels=tex[t][0], wrap=tex[t][1]), strategy=1,
                            bounds=[-0.05, -0.05, -0.05, 1.05, 1.05, 1.05])
        parts.append({tuple(p[:3]): p[3] for p in r["voxels"].tolist()})
    only0 = set(parts[0]) - set(parts[1])
    only1 = set(parts[1]) - set(parts[0])
    got_map = {tuple(p[:3]): p[3] for p in o2v.sort_voxels(got).tolist()}
    assert set(got_map) == set(parts[0]) | set(parts[1])
    assert all(got_map[k] == parts[0][k] for k in only0) and all(got_map[k] == parts[1][k] for k in only1)
    assert len(only0) > 100 and len(only1) > 100
