"""TEST INFRASTRUCTURE ONLY — ctypes binding of oracle/_ref/libo2vref*.so (the UNMODIFIED reference, built by
oracle/Makefile from /root/reference).  Importers: tests/, tests/golden/make_golden.py, bench.py's reference arm.
Never imported by the product package obj2voxel_b200.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def lib_path(patched_downscale=False):
    return os.path.join(_HERE, "_ref", "libo2vref_ss.so" if patched_downscale else "libo2vref.so")


def available(patched_downscale=False):
    return os.path.exists(lib_path(patched_downscale))


def _load(patched_downscale=False):
    key = bool(patched_downscale)
    if key in _LIBS:
        return _LIBS[key]
    # RTLD_LOCAL (default): the reference exports the same obj2voxel_* names as the product library
    lib = C.CDLL(lib_path(patched_downscale))
    fp = C.POINTER(C.c_float)
    u8p = C.POINTER(C.c_uint8)
    ip = C.POINTER(C.c_int)
    lib.o2vref_run_api.restype = C.c_longlong
    lib.o2vref_run_api.argtypes = [fp, fp, C.c_size_t, u8p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int,
                                   C.c_uint32, C.c_uint32, C.c_int, fp, ip, C.c_int, C.c_int,
                                   C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_double), C.POINTER(C.c_size_t)]
    lib.o2vref_run_internal.restype = C.c_longlong
    lib.o2vref_run_internal.argtypes = [fp, fp, u8p, fp, C.c_size_t, u8p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int,
                                        C.c_uint32, C.c_uint32, C.c_int, fp, ip, C.c_int,
                                        C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(fp), fp]
    lib.o2vref_read_voxel_file.restype = C.c_longlong
    lib.o2vref_read_voxel_file.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.POINTER(C.c_uint32))]
    lib.o2vref_run_file.restype = C.c_longlong
    lib.o2vref_run_file.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                    C.POINTER(C.POINTER(C.c_uint32))]
    lib.o2vref_free.argtypes = [C.c_void_p]
    lib.o2vref_hardware_threads.restype = C.c_uint
    _LIBS[key] = lib
    return lib


def _fptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _prep(verts, uvs, texture, bounds, unit):
    verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 9)
    uvs = None if uvs is None else np.ascontiguousarray(uvs, dtype=np.float32).reshape(-1, 6)
    tex = (None, 0, 0, 0, 1)
    keep = None
    if texture is not None:
        pixels = np.ascontiguousarray(texture["pixels"], dtype=np.uint8)
        h, w, ch = pixels.shape
        keep = pixels
        tex = (pixels.ctypes.data_as(C.POINTER(C.c_uint8)), w, h, ch, int(texture.get("wrap", 1)))
    b = None if bounds is None else np.ascontiguousarray(bounds, dtype=np.float32)
    u = None if unit is None else np.ascontiguousarray(unit, dtype=np.int32)
    return verts, uvs, tex, b, u, keep


def hardware_threads():
    return int(_load().o2vref_hardware_threads())


def run_api(verts, resolution, uvs=None, texture=None, supersampling=1, strategy=0, bounds=None, unit=None,
            workers=0, collect=True, patched_downscale=False):
    """Reference public C API.  Returns dict(voxels=(n,4) u32 sorted by (x,y,z) or None, count, seconds, sink_calls)."""
    lib = _load(patched_downscale)
    verts, uvs, tex, b, u, _keep = _prep(verts, uvs, texture, bounds, unit)
    out = C.POINTER(C.c_uint32)()
    secs = C.c_double()
    calls = C.c_size_t()
    n = lib.o2vref_run_api(_fptr(verts), _fptr(uvs), len(verts), tex[0], tex[1], tex[2], tex[3], tex[4],
                           resolution, supersampling, strategy, _fptr(b),
                           None if u is None else u.ctypes.data_as(C.POINTER(C.c_int)),
                           workers, 1 if collect else 0, C.byref(out), C.byref(secs), C.byref(calls))
    if n < 0:
        raise RuntimeError("reference returned error code %d" % (-n))
    voxels = None
    if collect:
        voxels = np.ctypeslib.as_array(out, shape=(max(n, 1), 4))[:n].copy()
        lib.o2vref_free(out)
        voxels = sort_voxels(voxels)
    return dict(voxels=voxels, count=int(n), seconds=secs.value, sink_calls=int(calls.value))


def run_internal(verts, resolution, uvs=None, types=None, colors=None, texture=None, supersampling=1, strategy=0,
                 bounds=None, unit=None, apply_downscale=False, patched_downscale=False):
    """Reference internal pipeline; float weights/colours.  Returns dict(xyz (n,3) u32, wrgb (n,4) f32, transform (12,))
    sorted by (x,y,z)."""
    lib = _load(patched_downscale)
    verts, uvs, tex, b, u, _keep = _prep(verts, uvs, texture, bounds, unit)
    t = None if types is None else np.ascontiguousarray(types, dtype=np.uint8)
    c = None if colors is None else np.ascontiguousarray(colors, dtype=np.float32).reshape(-1, 3)
    xyz = C.POINTER(C.c_uint32)()
    wrgb = C.POINTER(C.c_float)()
    transform = np.zeros(12, dtype=np.float32)
    n = lib.o2vref_run_internal(_fptr(verts), _fptr(uvs),
                                None if t is None else t.ctypes.data_as(C.POINTER(C.c_uint8)), _fptr(c), len(verts),
                                tex[0], tex[1], tex[2], tex[3], tex[4], resolution, supersampling, strategy, _fptr(b),
                                None if u is None else u.ctypes.data_as(C.POINTER(C.c_int)),
                                1 if apply_downscale else 0, C.byref(xyz), C.byref(wrgb), _fptr(transform))
    x = np.ctypeslib.as_array(xyz, shape=(max(n, 1), 3))[:n].copy()
    w = np.ctypeslib.as_array(wrgb, shape=(max(n, 1), 4))[:n].copy()
    lib.o2vref_free(xyz)
    lib.o2vref_free(wrgb)
    order = np.lexsort((x[:, 2], x[:, 1], x[:, 0]))
    return dict(xyz=x[order], wrgb=w[order], transform=transform)


def read_voxel_file(path, type_):
    """Reads a voxel file with the reference's own voxelio reader ("qef", "vox", "vl32"); (n, 4) u32 sorted by (x,y,z)."""
    lib = _load()
    out = C.POINTER(C.c_uint32)()
    n = lib.o2vref_read_voxel_file(str(path).encode(), type_.encode(), C.byref(out))
    if n < 0:
        raise RuntimeError("voxelio reader failed for %s (%d)" % (path, n))
    voxels = np.ctypeslib.as_array(out, shape=(max(n, 1), 4))[:n].copy()
    lib.o2vref_free(out)
    return sort_voxels(voxels)


def run_file(input_path, resolution, output_path=None, supersampling=1, strategy=0, workers=0):
    """The reference from an input file (its own OBJ / STL readers) to sorted voxels, or to an output file written by its
    own writers (then returns None)."""
    lib = _load()
    out = C.POINTER(C.c_uint32)()
    n = lib.o2vref_run_file(str(input_path).encode(), None if output_path is None else str(output_path).encode(),
                            resolution, supersampling, strategy, workers, C.byref(out))
    if n < 0:
        raise RuntimeError("reference returned error code %d" % (-n))
    if output_path is not None:
        return None
    voxels = np.ctypeslib.as_array(out, shape=(max(n, 1), 4))[:n].copy()
    lib.o2vref_free(out)
    return sort_voxels(voxels)


def sort_voxels(voxels):
    """Canonical order for comparisons: ascending (x, y, z)."""
    v = np.asarray(voxels).reshape(-1, 4)
    order = np.lexsort((v[:, 2], v[:, 1], v[:, 0]))
    return v[order]
