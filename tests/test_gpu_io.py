"""GPU (-m gpu): the I/O adapters either side of the hot path through the reference API:
binary STL and OBJ(+MTL) input files, VL32 / PLY / XYZRGB output (file and memory)."""
import struct

import numpy as np
import pytest

import obj2voxel_b200 as o2v
from obj2voxel_b200 import _lib, meshes
from oracle import oracle

pytestmark = pytest.mark.gpu


def write_binary_stl(path, tris):
    with open(path, "wb") as f:
        f.write(b"obj2voxel_b200 test".ljust(80, b" "))
        f.write(struct.pack("<I", len(tris)))
        for t in tris:
            f.write(struct.pack("<12f", 0.0, 0.0, 0.0, *[float(x) for x in t]))
            f.write(b"\0\0")


def run_file_job(in_path, resolution, out=None, out_type=None, strategy=0):
    inst = o2v.Instance()
    inst.set_input_file(in_path)
    if out is None:
        inst.set_output_callback()
    elif out == "memory":
        inst.set_output_memory(out_type)
    else:
        inst.set_output_file(out, out_type)
    inst.set_resolution(resolution)
    inst.set_color_strategy(strategy)
    err = inst.voxelize()
    result = dict(err=err, voxels=inst.collected(), memory=inst.get_output_memory() if out == "memory" else None)
    inst.free()
    return result


def test_stl_input_matches_callback_input(tmp_path):
    tris = meshes.lumpy_sphere(16, 17)
    stl = str(tmp_path / "sphere.stl")
    write_binary_stl(stl, tris)
    r = run_file_job(stl, 64)
    want = oracle.voxelize(tris, 64)["voxels"]
    assert r["err"] == o2v.ERR_OK and np.array_equal(r["voxels"], want)


def test_ascii_stl_and_missing_file_are_input_errors(tmp_path):
    o2v.load().obj2voxel_set_log_level(_lib.LOG_SILENT)
    bad = tmp_path / "ascii.stl"
    bad.write_text("solid x\nendsolid x\n" + " " * 100)
    assert run_file_job(str(bad), 16)["err"] == _lib.ERR_IO_OPEN_INPUT
    assert run_file_job(str(tmp_path / "nope.stl"), 16)["err"] == _lib.ERR_IO_OPEN_INPUT
    o2v.load().obj2voxel_set_log_level(_lib.LOG_INFO)


def test_vl32_ply_xyzrgb_outputs_agree(tmp_path):
    tris = meshes.unit_cube()
    stl = str(tmp_path / "cube.stl")
    write_binary_stl(stl, tris)
    want = oracle.voxelize(tris, 32)["voxels"]

    mem = run_file_job(stl, 32, out="memory", out_type="vl32")
    quads = o2v.sort_voxels(np.frombuffer(mem["memory"], dtype=">u4").reshape(-1, 4).astype(np.uint32))
    assert np.array_equal(quads, want)  # VL32: 16 big-endian bytes per voxel (voxelio vl32.cpp:83-88)

    vl32 = str(tmp_path / "cube.vl32")
    assert run_file_job(stl, 32, out=vl32)["err"] == o2v.ERR_OK
    assert open(vl32, "rb").read() == mem["memory"] or \
        np.array_equal(o2v.sort_voxels(np.fromfile(vl32, dtype=">u4").reshape(-1, 4).astype(np.uint32)), want)

    ply = str(tmp_path / "cube.ply")
    assert run_file_job(stl, 32, out=ply)["err"] == o2v.ERR_OK
    data = open(ply, "rb").read()
    head_end = data.index(b"end_header\r\n") + len(b"end_header\r\n")
    header = data[:head_end].decode()
    assert header.startswith("ply\r\nformat binary_big_endian 1.0\r\n")
    assert "element vertex %d\r\n" % len(want) in header  # count patched in on finalize (ply.cpp:63-77)
    body = o2v.sort_voxels(np.frombuffer(data[head_end:], dtype=">u4").reshape(-1, 4).astype(np.uint32))
    assert np.array_equal(body, want)  # PLY body is the VL32 body

    xyz = str(tmp_path / "cube.xyzrgb")
    assert run_file_job(stl, 32, out=xyz)["err"] == o2v.ERR_OK
    rows = np.loadtxt(xyz, dtype=np.int64).reshape(-1, 6)
    assert len(rows) == len(want) and np.all(rows[:, 3:] == 255)


def test_palette_formats_report_open_output_error(tmp_path):
    o2v.load().obj2voxel_set_log_level(_lib.LOG_SILENT)
    stl = str(tmp_path / "cube.stl")
    write_binary_stl(stl, meshes.unit_cube())
    assert run_file_job(stl, 16, out=str(tmp_path / "cube.qef"))["err"] == _lib.ERR_IO_OPEN_OUTPUT
    o2v.load().obj2voxel_set_log_level(_lib.LOG_INFO)


def test_obj_with_materials(tmp_path):
    """v / vt / f with fan triangulation and negative indices, mtllib + usemtl: Kd colours become UNTEXTURED triangles
    (reference src/io.cpp:301-302), faces without material stay white."""
    (tmp_path / "m.mtl").write_text("newmtl red\nKd 1 0 0\nnewmtl green\nKd 0 1 0\n")
    obj = tmp_path / "m.obj"
    obj.write_text("mtllib m.mtl\n"
                   "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\n"
                   "v 0 0 1\nv 1 0 1\nv 1 1 1\nv 0 1 1\n"
                   "usemtl red\nf 1 2 3 4\n"          # quad -> two triangles, z = 0 plane
                   "usemtl green\nf -4 -3 -2 -1\n"    # negative indices: z = 1 plane
                   "usemtl nothing\nf 1 5 8\n")       # unknown material -> no material -> white
    r = run_file_job(str(obj), 32, strategy=0)
    verts = np.array([[0, 0, 0, 1, 0, 0, 1, 1, 0], [0, 0, 0, 1, 1, 0, 0, 1, 0],
                      [0, 0, 1, 1, 0, 1, 1, 1, 1], [0, 0, 1, 1, 1, 1, 0, 1, 1],
                      [0, 0, 0, 0, 0, 1, 0, 1, 1]], dtype=np.float32)
    types = np.array([2, 2, 2, 2, 1], dtype=np.uint8)
    colors = np.array([[1, 0, 0], [1, 0, 0], [0, 1, 0], [0, 1, 0], [0, 0, 0]], dtype=np.float32)
    want = oracle.voxelize(verts, 32, types=types, colors=colors, strategy=0)["voxels"]
    assert r["err"] == o2v.ERR_OK and np.array_equal(r["voxels"], want)
    assert {0xFFFF0000, 0xFF00FF00, 0xFFFFFFFF} <= set(np.unique(r["voxels"][:, 3]).tolist())
