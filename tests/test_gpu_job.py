"""GPU (-m gpu): obj2voxel_voxelize()'s host-to-host job (obj2voxel_b200/csrc/o2v_job.cpp) — how the bytes travel must
not change a single record: bitmaps or records over PCIe, pinned or pageable input, one device or several (Z-slabs, the
triangles exchanged over peer memory).  Reference for every comparison: the CPU oracle."""
import numpy as np
import pytest

import obj2voxel_b200 as o2v
from obj2voxel_b200 import _lib, meshes
from oracle import oracle

pytestmark = pytest.mark.gpu


def run_bulk(verts, resolution, devices=None, uvs=None, texture=None, **kw):
    inst = o2v.Instance()
    inst.set_input_triangles(verts, uvs=uvs, texture=texture)
    inst.set_output_callback()
    inst.set_resolution(resolution)
    inst.set_color_strategy(kw.get("strategy", o2v.MAX_STRATEGY))
    inst.set_supersampling(kw.get("supersampling", 1))
    if kw.get("bounds") is not None:
        inst.set_mesh_boundaries(kw["bounds"])
    if devices is not None:
        inst.set_devices(devices)
    err = inst.voxelize()
    voxels, stats = inst.collected(), inst.stats()
    inst.free()
    assert err == o2v.ERR_OK
    return voxels, stats


def device_count():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("supersampling", [1, 2])
@pytest.mark.parametrize("parts", [1, 3])
def test_every_download_format_delivers_the_same_records(monkeypatch, parts, supersampling):
    """One GPU: an all-white result leaves the device as packed positions (the default: 4 bytes per voxel) or as occupancy
    bitmaps (1 bit per output voxel of the touched chunks) that host threads expand into the quads the callback receives;
    O2V_B200_DOWNLOAD=records forces the 16-byte records through PCIe instead.  Same records every way, and the
    oracle's."""
    verts = meshes.random_triangles(6000, 0.02, seed=41)
    res = 160 // supersampling
    monkeypatch.setenv("O2V_B200_PIPELINE_PARTS", str(parts))
    monkeypatch.setenv("O2V_B200_DOWNLOAD", "bitmap")
    bitmaps, stats = run_bulk(verts, res, bounds=meshes.UNIT_BOUNDS, supersampling=supersampling)
    assert stats["occupancy_path"] == 1
    monkeypatch.setenv("O2V_B200_DOWNLOAD", "records")
    records, stats = run_bulk(verts, res, bounds=meshes.UNIT_BOUNDS, supersampling=supersampling)
    assert np.array_equal(bitmaps, records) and stats["download_bytes"] == 16 * len(records)
    monkeypatch.delenv("O2V_B200_DOWNLOAD")
    packed, stats = run_bulk(verts, res, bounds=meshes.UNIT_BOUNDS, supersampling=supersampling)
    assert np.array_equal(packed, records) and stats["download_bytes"] == 4 * len(records)
    want = oracle.voxelize(verts, res, bounds=meshes.UNIT_BOUNDS, supersampling=supersampling)["voxels"]
    assert np.array_equal(bitmaps, want)


def test_packed_positions_of_a_grid_beyond_ten_bits():
    """Output grids above 1024 per axis take 8 bytes per position."""
    verts = meshes.random_triangles(2000, 0.01, seed=46)
    got, stats = run_bulk(verts, 1100, bounds=meshes.UNIT_BOUNDS)
    want = oracle.voxelize(verts, 1100, bounds=meshes.UNIT_BOUNDS)["voxels"]
    assert np.array_equal(got, want) and stats["download_bytes"] == 8 * len(want)


def test_bitmap_download_of_a_slab_and_of_an_unaligned_grid(monkeypatch):
    """Resolution 100 (chunk grid 128, voxels beyond the resolution kept like the reference keeps them) and a job
    restricted to a Z-slab: the bitmaps' chunk geometry must survive the download."""
    monkeypatch.setenv("O2V_B200_DOWNLOAD", "bitmap")
    verts = meshes.random_triangles(3000, 0.05, seed=42)
    got, _ = run_bulk(verts, 100, bounds=meshes.UNIT_BOUNDS)
    want = oracle.voxelize(verts, 100, bounds=meshes.UNIT_BOUNDS)["voxels"]
    assert np.array_equal(got, want)
    inst = o2v.Instance()
    inst.set_input_triangles(verts)
    inst.set_output_callback()
    inst.set_resolution(100)
    inst.set_mesh_boundaries(meshes.UNIT_BOUNDS)
    inst.set_slab(40, 72)
    assert inst.voxelize() == o2v.ERR_OK
    slab = inst.collected()
    inst.free()
    assert np.array_equal(slab, want[(want[:, 2] >= 40) & (want[:, 2] < 72)])


def test_pageable_input_is_staged_by_host_threads(monkeypatch):
    """A plain numpy array (pageable memory) above the staging threshold goes to the device in 4 MiB pieces through the
    host threads' pinned buffers; a pinned copy of the same array goes as it is.  Same records."""
    import torch

    n = 300_000  # 10.8 MB of vertices: above 2 staging chunks
    verts = meshes.random_triangles(n, 0.004, seed=43)
    pageable, _ = run_bulk(verts, 256, bounds=meshes.UNIT_BOUNDS)
    pinned = torch.from_numpy(verts).pin_memory()
    direct, _ = run_bulk(pinned.numpy(), 256, bounds=meshes.UNIT_BOUNDS)
    assert len(pageable) > 100_000
    assert np.array_equal(pageable, direct)
    engine = o2v.Engine(0)
    dev, _ = engine.voxelize_host(verts, o2v.make_params(resolution=256, bounds=meshes.UNIT_BOUNDS))
    engine.close()
    assert np.array_equal(pageable, o2v.sort_voxels(dev))


@pytest.mark.parametrize("supersampling", [1, 2])
@pytest.mark.parametrize("pieces", [2, 4])
def test_a_pinned_mesh_is_voxelized_piece_by_piece_while_it_uploads(monkeypatch, pieces, supersampling):
    """One GPU, an all-white mesh with known bounds in pinned memory: the triangle array goes up in pieces and every piece
    is voxelized while the next one crosses PCIe; a piece delivers the voxels no earlier piece has (o2v_b200_params::
    accumulate).  Same records as the job with the upload first (O2V_B200_STREAM_UPLOAD=0), and the oracle's."""
    import torch

    verts = meshes.random_triangles(40_000, 0.02, seed=47)
    pinned = torch.from_numpy(verts).pin_memory().numpy()
    res = 192 // supersampling
    monkeypatch.setenv("O2V_B200_PIPELINE_PARTS", str(pieces))
    monkeypatch.setenv("O2V_B200_STREAM_MIN", "1000")
    monkeypatch.setenv("O2V_B200_STREAM_UPLOAD", "0")
    first, stats = run_bulk(pinned, res, bounds=meshes.UNIT_BOUNDS, supersampling=supersampling)
    monkeypatch.setenv("O2V_B200_STREAM_UPLOAD", "1")
    lib = o2v.load()
    lines = []
    def on_log(data, message, level):
        lines.append(message.decode())
        return True

    callback = _lib.LOG_CALLBACK(on_log)
    lib.obj2voxel_set_log_callback(callback, None)
    lib.obj2voxel_set_log_level(_lib.LOG_DEBUG)
    try:
        streamed, streamed_stats = run_bulk(pinned, res, bounds=meshes.UNIT_BOUNDS, supersampling=supersampling)
    finally:
        lib.obj2voxel_set_log_callback(None, None)
        lib.obj2voxel_set_log_level(_lib.LOG_INFO)
    assert any("the upload runs under the parts" in line for line in lines), lines[-3:]
    assert np.array_equal(streamed, first) and len(streamed) == streamed_stats["voxels"]
    want = oracle.voxelize(verts, res, bounds=meshes.UNIT_BOUNDS, supersampling=supersampling)["voxels"]
    assert np.array_equal(streamed, want)
    # an ordinary job afterwards starts from clean bitmaps
    again, _ = run_bulk(verts[:5000], res, bounds=meshes.UNIT_BOUNDS, supersampling=supersampling)
    assert np.array_equal(again, oracle.voxelize(verts[:5000], res, bounds=meshes.UNIT_BOUNDS,
                                                 supersampling=supersampling)["voxels"])


@pytest.mark.parametrize("case", ["white", "white_autobounds", "white_ss2", "white_skewed", "textured_blend"])
def test_two_devices_deliver_the_oracles_records(case):
    """Two GPUs of one process: Z-slabs of whole chunk rows, each device uploads half of the triangles and stores every
    triangle into the memory of the device(s) whose slab it can reach (all-white meshes), or takes the whole mesh
    (coloured ones: the fold replays the triangle order).  The union of the slabs is the oracle's result."""
    if device_count() < 2:
        pytest.skip("needs two CUDA devices")
    verts = meshes.random_triangles(20_000, 0.03, seed=44)
    kw, okw, extra = dict(bounds=meshes.UNIT_BOUNDS), dict(bounds=meshes.UNIT_BOUNDS), {}
    res = 256
    if case == "white_autobounds":
        kw, okw = {}, {}
    elif case == "white_skewed":
        # nine tenths of the triangles below z = 0.2: the slabs are balanced by triangles per chunk row, not split in
        # the middle (and the union is still the oracle's result)
        low = meshes.random_triangles(18_000, 0.03, seed=47)
        low[:, 2::3] *= np.float32(0.2)
        verts = np.concatenate([low, meshes.random_triangles(2_000, 0.03, seed=48)])
    elif case == "white_ss2":
        res = 128
        kw.update(supersampling=2)
        okw.update(supersampling=2)
    elif case == "textured_blend":
        uvs = meshes.random_uvs(len(verts), seed=45)
        pixels = meshes.random_texture(32, 16, 3)
        kw.update(strategy=o2v.BLEND_STRATEGY)
        okw.update(strategy=o2v.BLEND_STRATEGY, uvs=uvs, texture=dict(pixels=pixels, wrap=o2v.UV_WRAP))
        extra = dict(uvs=uvs, texture=o2v.Texture(pixels, wrap=o2v.UV_WRAP))
    got, stats = run_bulk(verts, res, devices=[0, 1], **kw, **extra)
    want = oracle.voxelize(verts, res, **okw)["voxels"]
    assert np.array_equal(got, want)
    if case != "textured_blend":
        # the exchange left each device the triangles of its slab only: fewer than twice the mesh in total
        assert stats["occupancy_path"] == 1 and stats["slab_triangles"] < 1.2 * len(verts)
    one, _ = run_bulk(verts, res, devices=[0], **kw, **extra)
    assert np.array_equal(one, got)
