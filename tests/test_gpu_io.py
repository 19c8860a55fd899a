"""GPU (-m gpu): the I/O adapters either side of the hot path through the reference API:
binary STL and OBJ(+MTL) input files, VL32 / PLY / XYZRGB / QEF / VOX output (file and memory), and the CLI."""
import os
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


def parse_qef(path):
    """Qubicle Exchange Format as the reference writes it (voxelio/src/format/qef.cpp:212-295)."""
    lines = open(path).read().split("\n")
    assert lines[:3] == ["Qubicle Exchange Format", "Version 0.2", "www.minddesk.com"]
    dims = [int(x) for x in lines[3].split()]
    n_colors = int(lines[4])
    palette = []
    for line in lines[5:5 + n_colors]:
        parts = line.split()
        assert all(len(x.split(".")[1]) == 4 for x in parts)  # stringifyFractionRpad(channel, 255, 4)
        palette.append([float(x) for x in parts])
    rows = np.array([[int(x) for x in line.split()] for line in lines[5 + n_colors:] if line], dtype=np.int64)
    return dims, np.array(palette), rows


def parse_vox(path):
    """MagicaVoxel 150 as the reference writes it (voxelio/src/format/vox.cpp:828-1080): returns (voxels (n, 4) with
    the palette index in column 3, palette (256, 4) RGBA with entry i = index i + 1, number of models)."""
    data = open(path, "rb").read()
    assert data[:4] == b"VOX " and struct.unpack_from("<I", data, 4)[0] == 150
    assert data[8:12] == b"MAIN"
    self_size, child_size = struct.unpack_from("<II", data, 12)
    assert self_size == 0 and child_size == len(data) - 20
    pos, models, translations, palette = 20, [], [], None
    while pos < len(data):
        cid = data[pos:pos + 4]
        size, children = struct.unpack_from("<II", data, pos + 4)
        body = data[pos + 12:pos + 12 + size]
        if cid == b"SIZE":
            assert struct.unpack("<III", body) == (256, 256, 256)
        elif cid == b"XYZI":
            n = struct.unpack_from("<I", body, 0)[0]
            models.append(np.frombuffer(body[4:4 + 4 * n], dtype=np.uint8).reshape(-1, 4).astype(np.int64))
        elif cid == b"nTRN":
            node_id = struct.unpack_from("<I", body, 0)[0]
            tail = body[body.rindex(b"_t") + 2:]
            length = struct.unpack_from("<I", tail, 0)[0]
            if node_id >= 2:
                translations.append([int(x) for x in tail[4:4 + length].decode().split()])
        elif cid == b"RGBA":
            palette = np.frombuffer(body, dtype=np.uint8).reshape(256, 4)
        else:
            assert cid in (b"nGRP", b"nSHP"), cid
        pos += 12 + size + children
    assert len(models) == len(translations) and palette is not None
    voxels = []
    for m, t in zip(models, translations):
        world = m.copy()
        world[:, :3] += np.array(t) - 128  # the transform node carries the model centre
        voxels.append(world)
    return np.concatenate(voxels), palette, len(models)


def test_qef_and_vox_outputs(tmp_path):
    """Palette formats: a two-colour OBJ at resolution 300 (> 256: two VOX models per axis)."""
    (tmp_path / "m.mtl").write_text("newmtl red\nKd 1 0 0\nnewmtl teal\nKd 0 0.5 0.5\n")
    obj = tmp_path / "m.obj"
    obj.write_text("mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0 0 1\nv 1 0 1\n"
                   "usemtl red\nf 1 2 3 4\nusemtl teal\nf 1 2 6 5\n")
    res = 300
    want = run_file_job(str(obj), res)["voxels"]
    assert len(want) > 100000 and set(np.unique(want[:, 3]).tolist()) >= {0xFFFF0000, 0xFF007F7F}

    qef = str(tmp_path / "m.qef")
    assert run_file_job(str(obj), res, out=qef)["err"] == o2v.ERR_OK
    dims, palette, rows = parse_qef(qef)
    assert dims == [res, res, res] and len(rows) == len(want)
    rgb = (np.floor(palette[rows[:, 3]] * 255 + 0.5)).astype(np.int64)  # 4 truncated decimals round back exactly
    argb = 0xFF000000 | (rgb[:, 0] << 16) | (rgb[:, 1] << 8) | rgb[:, 2]
    got = o2v.sort_voxels(np.concatenate([rows[:, :3], argb[:, None]], axis=1).astype(np.uint32))
    assert np.array_equal(got, want)

    vox = str(tmp_path / "m.vox")
    assert run_file_job(str(obj), res, out=vox)["err"] == o2v.ERR_OK
    voxels, rgba, n_models = parse_vox(vox)
    assert n_models > 1 and voxels[:, 3].min() >= 1  # palette index 0 is reserved
    colour = rgba[voxels[:, 3] - 1].astype(np.int64)
    argb = (colour[:, 3] << 24) | (colour[:, 0] << 16) | (colour[:, 1] << 8) | colour[:, 2]
    got = o2v.sort_voxels(np.concatenate([voxels[:, :3], argb[:, None]], axis=1).astype(np.uint32))
    assert np.array_equal(got, want)


def test_vox_palette_reduction_keeps_positions(tmp_path):
    """More than 255 colours: positions stay exact, colours move to at most 255 representatives close to the originals."""
    n = 1500
    v = meshes.random_triangles(n, 0.04, seed=3)
    uv = meshes.random_uvs(n, seed=4)
    tex = o2v.Texture(meshes.random_texture(64, 64, 3, seed=9), wrap=o2v.UV_WRAP)

    def job(out):
        inst = o2v.Instance()
        inst.set_input_callback(v, uvs=uv, texture=tex)
        if out is None:
            inst.set_output_callback()
        else:
            inst.set_output_file(out, None)
        inst.set_resolution(64)
        inst.set_color_strategy(1)
        inst.set_mesh_boundaries([0, 0, 0, 1, 1, 1])
        err = inst.voxelize()
        voxels = inst.collected()
        inst.free()
        assert err == o2v.ERR_OK
        return voxels

    want = job(None)
    assert len(np.unique(want[:, 3])) > 255
    vox = str(tmp_path / "t.vox")
    job(vox)
    voxels, rgba, _ = parse_vox(vox)
    order = np.lexsort((voxels[:, 2], voxels[:, 1], voxels[:, 0]))
    voxels = voxels[order]
    assert np.array_equal(voxels[:, :3], want[:, :3].astype(np.int64))
    assert len(np.unique(voxels[:, 3])) <= 255
    colour = rgba[voxels[:, 3] - 1].astype(np.int64)
    orig = np.stack([(want[:, 3] >> 16) & 255, (want[:, 3] >> 8) & 255, want[:, 3] & 255], axis=1).astype(np.int64)
    assert np.abs(colour[:, :3] - orig).mean() < 24  # median cut of ~64^3 random colours into 255 boxes


def test_cli_matches_the_api(tmp_path):
    """obj2voxel-b200 with the reference CLI's flags (src/main.cpp:264-380): -r, -s, -u, -p, -o, exit status."""
    import subprocess

    from conftest import ROOT

    exe = os.path.join(ROOT, "obj2voxel_b200", "obj2voxel-b200")
    tris = meshes.lumpy_sphere(20, 21) * np.float32([1.0, 0.6, 0.3] * 3)
    stl = str(tmp_path / "s.stl")
    write_binary_stl(stl, tris)
    out = str(tmp_path / "s.vl32")
    r = subprocess.run([exe, stl, out, "-r", "48", "-s", "blend", "-u", "-p", "zXy", "-j", "2"], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = o2v.sort_voxels(np.fromfile(out, dtype=">u4").reshape(-1, 4).astype(np.uint32))
    # -p zXy: output x = input z, output y = -input x, output z = input y
    want = oracle.voxelize(tris, 48, strategy=1, supersampling=2, unit=[0, 0, 1, -1, 0, 0, 0, 1, 0])["voxels"]
    assert np.array_equal(got, want)
    # explicit output format on an extension-less path, long flags
    out2 = str(tmp_path / "noext")
    r = subprocess.run([exe, stl, out2, "--res=16", "-oxyzrgb"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and len(open(out2).read().splitlines()) == len(oracle.voxelize(tris, 16)["voxels"])
    # failures surface in the exit status (stated deviation from the reference, which always exits 0)
    r = subprocess.run([exe, str(tmp_path / "missing.stl"), out, "-r", "16"], capture_output=True, text=True, timeout=300)
    assert r.returncode == _lib.ERR_IO_OPEN_INPUT


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


# ---- against the reference's own readers, writers and a real model -------------------------------------------------

def reference_harness():
    from oracle import refharness

    if not refharness.available():
        pytest.skip("oracle/_ref (the compiled reference) is not on this machine")
    return refharness


def test_cornell_box_obj_matches_the_reference(tmp_path):
    """Real-mesh parity of the OBJ(+MTL) reader: tinyobjloader's cornell_box.obj — quads, several objects,
    `usemtl` per object, comments and blank lines with trailing spaces — at resolution 64 gives the unmodified
    reference's 25 574 voxels (17 639 white / 3 970 red / 3 965 green; fixture made by tests/golden/make_golden_models.py
    from oracle/_ref with its tinyobjloader-based reader)."""
    from conftest import GOLDEN_DIR

    g = np.load(os.path.join(GOLDEN_DIR, "models", "cornell_box_r64.npz"))
    (tmp_path / "cornell_box.obj").write_bytes(g["obj"].tobytes())
    (tmp_path / "cornell_box.mtl").write_bytes(g["mtl"].tobytes())
    r = run_file_job(str(tmp_path / "cornell_box.obj"), int(g["resolution"]))
    assert r["err"] == o2v.ERR_OK
    colours, counts = np.unique(r["voxels"][:, 3], return_counts=True)
    assert dict(zip(colours.tolist(), counts.tolist())) == {0xFFFFFFFF: 17639, 0xFFFF0000: 3970, 0xFF00FF00: 3965}
    assert np.array_equal(r["voxels"], g["voxels"])


def test_output_files_read_back_with_the_references_readers(tmp_path):
    """QEF, VOX and VL32 files written by this library, parsed by voxelio's own readers (oracle/_ref), give the voxels
    the callback sink receives; the PLY header is the reference's fixed 299 bytes (voxelio/src/format/ply.cpp:18-35,
    63-77)."""
    ref = reference_harness()
    (tmp_path / "m.mtl").write_text("newmtl red\nKd 1 0 0\nnewmtl teal\nKd 0 0.5 0.5\n")
    obj = tmp_path / "m.obj"
    obj.write_text("mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0 0 1\nv 1 0 1\n"
                   "usemtl red\nf 1 2 3 4\nusemtl teal\nf 1 2 6 5\n")
    res = 96
    want = run_file_job(str(obj), res)["voxels"]
    for ext in ("qef", "vox", "vl32"):
        path = str(tmp_path / ("m." + ext))
        assert run_file_job(str(obj), res, out=path)["err"] == o2v.ERR_OK
        got = ref.read_voxel_file(path, ext)
        assert np.array_equal(got, want), ext
    ply = str(tmp_path / "m.ply")
    assert run_file_job(str(obj), res, out=ply)["err"] == o2v.ERR_OK
    count = ("%d\r\ncomment " % len(want)).encode()
    placeholder = b"....;....;....;....;....;...\r\n"
    expected = (b"ply\r\nformat binary_big_endian 1.0\r\n"
                b"comment generated by voxel-io: a C++ library by Jan \"Eisenwave\" Schultke\r\n"
                b"element vertex " + count + placeholder[len(count):] +
                b"property int x\r\nproperty int y\r\nproperty int z\r\n"
                b"property uchar alpha\r\nproperty uchar red\r\nproperty uchar green\r\nproperty uchar blue\r\n"
                b"end_header\r\n")
    data = open(ply, "rb").read()
    assert len(expected) == 299 and data[:299] == expected  # (ply.cpp:18 says 300; what it writes is 299)
    assert len(data) == 299 + 16 * len(want)


def test_output_files_equal_the_references_own_files(tmp_path):
    """The same OBJ through the reference (oracle/_ref, its tinyobjloader reader and voxelio writers) and through this
    library: VL32, PLY and XYZRGB files hold the same records (the order of a voxel list is free), QEF and VOX read back
    to the same voxels.  (OBJ, not STL: the reference's STL stream mis-indexes its vertex array)"""
    ref = reference_harness()
    tris = meshes.lumpy_sphere(14, 15)
    stl = str(tmp_path / "s.obj")
    with open(stl, "w") as f:
        for t in tris:
            for k in range(3):
                f.write("v %r %r %r\n" % tuple(float(x) for x in t[3 * k:3 * k + 3]))
        for i in range(len(tris)):
            f.write("f %d %d %d\n" % (3 * i + 1, 3 * i + 2, 3 * i + 3))

    def body(path, skip):
        raw = open(path, "rb").read()[skip:]
        return o2v.sort_voxels(np.frombuffer(raw, dtype=">u4").reshape(-1, 4).astype(np.uint32))

    for ext, skip in (("vl32", 0), ("ply", 299)):
        ours, theirs = str(tmp_path / ("ours." + ext)), str(tmp_path / ("theirs." + ext))
        assert run_file_job(stl, 48, out=ours)["err"] == o2v.ERR_OK
        ref.run_file(stl, 48, output_path=theirs)
        assert os.path.getsize(ours) == os.path.getsize(theirs)
        assert np.array_equal(body(ours, skip), body(theirs, skip)), ext
        if ext == "ply":
            assert open(ours, "rb").read()[:299] == open(theirs, "rb").read()[:299]
    ours, theirs = str(tmp_path / "ours.xyzrgb"), str(tmp_path / "theirs.xyzrgb")
    assert run_file_job(stl, 48, out=ours)["err"] == o2v.ERR_OK
    ref.run_file(stl, 48, output_path=theirs)
    assert sorted(open(ours).read().splitlines()) == sorted(open(theirs).read().splitlines())
    for ext in ("qef", "vox"):
        ours, theirs = str(tmp_path / ("ours." + ext)), str(tmp_path / ("theirs." + ext))
        assert run_file_job(stl, 48, out=ours)["err"] == o2v.ERR_OK
        ref.run_file(stl, 48, output_path=theirs)
        assert np.array_equal(ref.read_voxel_file(ours, ext), ref.read_voxel_file(theirs, ext)), ext


def test_obj_face_longer_than_any_line_buffer(tmp_path):
    """A 600-gon on one `f` line (about 8 KB, above the 4 KB a fixed line buffer used to hold) with long-hand indices:
    every vertex is read and the polygon is fan-triangulated."""
    n = 600
    angle = np.linspace(0.0, 2.0 * np.pi, n, endpoint=False)
    ring = np.stack([0.5 + 0.45 * np.cos(angle), 0.5 + 0.45 * np.sin(angle), np.full(n, 0.25)], axis=1).astype(np.float32)
    obj = tmp_path / "ngon.obj"
    with open(obj, "w") as f:
        f.write("v 0 0 0\nv 1 1 1\n")  # bounds
        for p in ring:
            f.write("v %r %r %r\n" % tuple(float(x) for x in p))
        f.write("f " + " ".join("%d//%d" % (k + 3, k + 3) for k in range(n)) + "\n")
        f.write("f 1 2 2\n")  # a degenerate face keeps the bounds vertices referenced
    assert os.path.getsize(obj) > 8000
    r = run_file_job(str(obj), 64)
    tris = np.array([np.concatenate([ring[0], ring[k], ring[k + 1]]) for k in range(1, n - 1)], dtype=np.float32)
    assert r["err"] == o2v.ERR_OK
    # the mesh bounds come from every triangle, the degenerate one included
    full = np.concatenate([tris, np.array([[0, 0, 0, 1, 1, 1, 1, 1, 1]], dtype=np.float32)])
    want = oracle.voxelize(full, 64)["voxels"]
    assert np.array_equal(r["voxels"], want)
