"""TEST INFRASTRUCTURE ONLY — ctypes binding of oracle/liboracle.so (the plain-C restatement, oracle/o2v_oracle.c).
Importers: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg.  Never imported by obj2voxel_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MATERIALLESS, UNTEXTURED, TEXTURED = 1, 2, 3
MAX, BLEND = 0, 1
UV_CLAMP, UV_WRAP = 0, 1


class Texture(C.Structure):
    _fields_ = [("pixels", C.POINTER(C.c_uint8)), ("width", C.c_size_t), ("height", C.c_size_t),
                ("channels", C.c_int), ("wrap", C.c_int)]


class Params(C.Structure):
    _fields_ = [("resolution", C.c_uint32), ("supersampling", C.c_uint32), ("strategy", C.c_int),
                ("bounds_known", C.c_int), ("bounds", C.c_float * 6), ("unit_transform", C.c_int * 9),
                ("downscale", C.c_int), ("threads", C.c_int)]


class Result(C.Structure):
    _fields_ = [("count", C.c_size_t), ("xyz", C.POINTER(C.c_uint32)), ("argb", C.POINTER(C.c_uint32)),
                ("wrgb", C.POINTER(C.c_float)), ("transform", C.c_float * 12), ("contributions", C.c_uint64),
                ("subtriangles", C.c_uint64)]


def build():
    """Compile liboracle.so (gcc only; no reference sources needed)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        fp = C.POINTER(C.c_float)
        _LIB.o2v_oracle_voxelize.restype = C.c_int
        _LIB.o2v_oracle_voxelize.argtypes = [C.POINTER(Params), C.c_size_t, fp, fp, C.POINTER(C.c_uint8), fp,
                                             C.POINTER(Texture), C.POINTER(Result)]
        _LIB.o2v_oracle_free_result.argtypes = [C.POINTER(Result)]
        _LIB.o2v_oracle_ileave3.restype = C.c_uint64
        _LIB.o2v_oracle_ileave3.argtypes = [C.c_uint32] * 3
        _LIB.o2v_oracle_dileave3.argtypes = [C.c_uint64, C.POINTER(C.c_uint32)]
        _LIB.o2v_oracle_mesh_transform.argtypes = [fp, fp, C.c_uint32, C.POINTER(C.c_int), fp]
        _LIB.o2v_oracle_split.restype = C.c_int
        _LIB.o2v_oracle_split.argtypes = [C.c_uint32, C.c_uint32, fp, C.c_int, fp]
        _LIB.o2v_oracle_clip_voxel.restype = C.c_int
        _LIB.o2v_oracle_clip_voxel.argtypes = [fp, C.POINTER(C.c_uint32), C.c_float, fp]
        _LIB.o2v_oracle_subdivide.restype = C.c_size_t
        _LIB.o2v_oracle_subdivide.argtypes = [fp, fp, C.c_size_t]
        _LIB.o2v_oracle_quantize_argb.restype = C.c_uint32
        _LIB.o2v_oracle_quantize_argb.argtypes = [fp]
        _LIB.o2v_oracle_texture_lookup.argtypes = [C.POINTER(Texture), fp, fp]
    return _LIB


def _fptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def make_texture(texture):
    if texture is None:
        return None, None
    pixels = np.ascontiguousarray(texture["pixels"], dtype=np.uint8)
    h, w, ch = pixels.shape
    t = Texture(pixels.ctypes.data_as(C.POINTER(C.c_uint8)), w, h, ch, int(texture.get("wrap", UV_WRAP)))
    return t, pixels


def voxelize(verts, resolution, uvs=None, types=None, colors=None, texture=None, supersampling=1, strategy=MAX,
             bounds=None, unit=None, downscale=True, threads=0):
    """Returns dict(xyz (n,3) u32, argb (n,) u32, wrgb (n,4) f32, voxels (n,4) u32, transform, contributions,
    subtriangles), sorted ascending by (x, y, z)."""
    L = lib()
    verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 9)
    uvs = None if uvs is None else np.ascontiguousarray(uvs, dtype=np.float32).reshape(-1, 6)
    types = None if types is None else np.ascontiguousarray(types, dtype=np.uint8)
    colors = None if colors is None else np.ascontiguousarray(colors, dtype=np.float32).reshape(-1, 3)
    p = Params()
    p.resolution = resolution
    p.supersampling = supersampling
    p.strategy = strategy
    p.bounds_known = 0 if bounds is None else 1
    if bounds is not None:
        p.bounds = (C.c_float * 6)(*[float(b) for b in bounds])
    p.unit_transform = (C.c_int * 9)(*(unit if unit is not None else [1, 0, 0, 0, 1, 0, 0, 0, 1]))
    p.downscale = 1 if downscale else 0
    p.threads = threads
    tex, _keep = make_texture(texture)
    r = Result()
    err = L.o2v_oracle_voxelize(C.byref(p), len(verts), _fptr(verts), _fptr(uvs),
                                None if types is None else types.ctypes.data_as(C.POINTER(C.c_uint8)), _fptr(colors),
                                None if tex is None else C.byref(tex), C.byref(r))
    if err != 0:
        raise RuntimeError("oracle error %d" % err)
    n = r.count
    if n == 0:
        xyz, argb, wrgb = np.zeros((0, 3), np.uint32), np.zeros((0,), np.uint32), np.zeros((0, 4), np.float32)
    else:
        xyz = np.ctypeslib.as_array(r.xyz, shape=(n, 3)).copy()
        argb = np.ctypeslib.as_array(r.argb, shape=(n,)).copy()
        wrgb = np.ctypeslib.as_array(r.wrgb, shape=(n, 4)).copy()
    result = dict(transform=np.array(list(r.transform), dtype=np.float32), contributions=int(r.contributions),
                  subtriangles=int(r.subtriangles))
    L.o2v_oracle_free_result(C.byref(r))
    order = np.lexsort((xyz[:, 2], xyz[:, 1], xyz[:, 0]))
    result.update(xyz=xyz[order], argb=argb[order], wrgb=wrgb[order],
                  voxels=np.concatenate([xyz[order], argb[order][:, None]], axis=1).astype(np.uint32))
    return result


def ileave3(x, y, z):
    return int(lib().o2v_oracle_ileave3(x, y, z))


def dileave3(n):
    out = (C.c_uint32 * 3)()
    lib().o2v_oracle_dileave3(n, out)
    return tuple(out)


def mesh_transform(mesh_min, mesh_max, sample_resolution, unit=None):
    mn = np.asarray(mesh_min, dtype=np.float32)
    mx = np.asarray(mesh_max, dtype=np.float32)
    u = (C.c_int * 9)(*(unit if unit is not None else [1, 0, 0, 0, 1, 0, 0, 0, 1]))
    out = np.zeros(12, dtype=np.float32)
    lib().o2v_oracle_mesh_transform(_fptr(mn), _fptr(mx), sample_resolution, u, _fptr(out))
    return out


def split(axis, plane, tri15, keep_hi):
    t = np.ascontiguousarray(tri15, dtype=np.float32)
    out = np.zeros(45, dtype=np.float32)
    n = lib().o2v_oracle_split(axis, plane, _fptr(t), 1 if keep_hi else 0, _fptr(out))
    return out.reshape(3, 15)[:n].copy()


def clip_voxel(tri15, pos, whole_area):
    t = np.ascontiguousarray(tri15, dtype=np.float32)
    p = (C.c_uint32 * 3)(*pos)
    out = np.zeros(3, dtype=np.float32)
    n = lib().o2v_oracle_clip_voxel(_fptr(t), p, float(whole_area), _fptr(out))
    return n, out


def subdivide(tri15, cap=1 << 16):
    t = np.ascontiguousarray(tri15, dtype=np.float32)
    out = np.zeros((cap, 15), dtype=np.float32)
    n = lib().o2v_oracle_subdivide(_fptr(t), _fptr(out), cap)
    if n > cap:
        return subdivide(tri15, cap=n)
    return out[:n].copy()


def quantize_argb(rgb):
    c = np.ascontiguousarray(rgb, dtype=np.float32)
    return int(lib().o2v_oracle_quantize_argb(_fptr(c)))
