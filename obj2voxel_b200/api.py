"""Host-side mirror of the reference's plugin interface for the voxelization path.

Two layers, both thin wrappers over the C ABI (include/*.h) — no voxel is ever computed in Python:

* `Instance` / `Texture`: the reference's own API (include/obj2voxel.h), same names, argument meaning and error codes, so
  the parity tests read like reference test/main.cpp.
* `Engine`: the additive bulk/device API (include/obj2voxel_b200.h) used by bench.py and the multi-GPU slab driver;
  accepts numpy arrays (host path: H2D + kernels + D2H) or torch CUDA tensors (device-resident path).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (BLEND_STRATEGY, ERR_OK, MAX_STRATEGY, UV_CLAMP, UV_WRAP, Mesh, Params, Stats)  # noqa: F401


class DeviceError(RuntimeError):
    """The CUDA path is unavailable or failed.  There is no CPU fallback."""


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Texture:
    """obj2voxel_texture (include/obj2voxel.h: texture_alloc / load_pixels / teture_set_uv_mode / texture_free)."""

    def __init__(self, pixels=None, wrap=UV_WRAP):
        self._lib = _lib.load()
        self.handle = self._lib.obj2voxel_texture_alloc()
        if pixels is not None:
            self.load_pixels(pixels)
            self.set_uv_mode(wrap)

    def load_pixels(self, pixels):
        p = np.ascontiguousarray(pixels, dtype=np.uint8)
        h, w, ch = p.shape
        ok = self._lib.obj2voxel_texture_load_pixels(self.handle, p.ctypes.data_as(C.c_char_p), w, h, ch)
        if not ok:
            raise ValueError("texture_load_pixels failed")

    def load_from_memory(self, data, type_="png"):
        return bool(self._lib.obj2voxel_texture_load_from_memory(self.handle, data, len(data), type_.encode()))

    def load_from_file(self, path, type_=None):
        return bool(self._lib.obj2voxel_texture_load_from_file(self.handle, path.encode(),
                                                               None if type_ is None else type_.encode()))

    def set_uv_mode(self, mode):
        self._lib.obj2voxel_teture_set_uv_mode(self.handle, mode)  # sic

    def meta(self):
        w, h, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        self._lib.obj2voxel_texture_get_meta(self.handle, C.byref(w), C.byref(h), C.byref(c))
        return w.value, h.value, c.value

    def pixels(self):
        w, h, c = self.meta()
        buf = C.create_string_buffer(w * h * c)
        self._lib.obj2voxel_texture_get_pixels(self.handle, buf)
        return np.frombuffer(buf.raw, dtype=np.uint8).reshape(h, w, c).copy()

    def free(self):
        if self.handle:
            self._lib.obj2voxel_texture_free(self.handle)
            self.handle = None


class Instance:
    """obj2voxel_instance with the reference's setters; `voxelize()` returns the reference's error code."""

    def __init__(self):
        self._lib = _lib.load()
        self.handle = self._lib.obj2voxel_alloc()
        self._keep = []
        self.voxels = []  # filled by collecting output callbacks
        self.sink_calls = 0

    # -- configuration ------------------------------------------------------------------------------------------
    def set_resolution(self, r):
        self._lib.obj2voxel_set_resolution(self.handle, r)

    def set_supersampling(self, level):
        self._lib.obj2voxel_set_supersampling(self.handle, level)

    def set_color_strategy(self, s):
        self._lib.obj2voxel_set_color_strategy(self.handle, s)

    def set_texture(self, texture):
        self._lib.obj2voxel_set_texture(self.handle, texture.handle)

    def set_parallel(self, enabled):
        self._lib.obj2voxel_set_parallel(self.handle, enabled)

    def set_unit_transform(self, t9):
        arr = (C.c_int * 9)(*[int(x) for x in t9])
        self._lib.obj2voxel_set_unit_transform(self.handle, arr)

    def set_mesh_boundaries(self, b6):
        arr = (C.c_float * 6)(*[float(x) for x in b6])
        self._lib.obj2voxel_set_mesh_boundaries(self.handle, arr)

    def get_resolution(self):
        return self._lib.obj2voxel_get_resolution(self.handle)

    def get_chunk_size(self):
        return self._lib.obj2voxel_get_chunk_size(self.handle)

    # -- input --------------------------------------------------------------------------------------------------
    def set_input_callback(self, verts, uvs=None, texture=None, colors=None):
        """Feeds triangles one per callback invocation through obj2voxel_set_triangle_* like reference test inputs
        (test/testutil.hpp TriangleInput)."""
        verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 9)
        uvs = None if uvs is None else np.ascontiguousarray(uvs, dtype=np.float32).reshape(-1, 6)
        colors = None if colors is None else np.ascontiguousarray(colors, dtype=np.float32).reshape(-1, 3)
        state = {"i": 0}
        lib = self._lib

        def callback(_data, out_triangle):
            i = state["i"]
            if i >= len(verts):
                return False
            if uvs is not None and texture is not None:
                lib.obj2voxel_set_triangle_textured(out_triangle, _fptr(verts[i]), _fptr(uvs[i]), texture.handle)
            elif colors is not None:
                lib.obj2voxel_set_triangle_colored(out_triangle, _fptr(verts[i]), _fptr(colors[i]))
            else:
                lib.obj2voxel_set_triangle_basic(out_triangle, _fptr(verts[i]))
            state["i"] = i + 1
            return True

        cb = _lib.TRIANGLE_CALLBACK(callback)
        self._keep += [cb, verts, uvs, colors, texture]
        lib.obj2voxel_set_input_callback(self.handle, cb, None)

    def set_input_triangles(self, verts, uvs=None, texture=None):
        """Additive bulk input (obj2voxel_b200_set_input_triangles)."""
        verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 9)
        uvs = None if uvs is None else np.ascontiguousarray(uvs, dtype=np.float32).reshape(-1, 6)
        self._keep += [verts, uvs, texture]
        self._lib.obj2voxel_b200_set_input_triangles(self.handle, _fptr(verts), None if uvs is None else _fptr(uvs),
                                                     len(verts), None if texture is None else texture.handle)

    def set_input_file(self, path, type_=None):
        p = path.encode()
        self._keep.append(p)
        self._lib.obj2voxel_set_input_file(self.handle, p, None if type_ is None else type_.encode())

    # -- output -------------------------------------------------------------------------------------------------
    def set_output_callback(self, fail_after=None):
        """Collecting sink (test/testutil.hpp CountingOutput / VoxelioOutput analogue)."""
        def callback(_data, quads, count):
            self.sink_calls += 1
            if fail_after is not None and self.sink_calls > fail_after:
                return False
            if count:
                self.voxels.append(np.ctypeslib.as_array(quads, shape=(count, 4)).copy())
            return True

        cb = _lib.VOXEL_CALLBACK(callback)
        self._keep.append(cb)
        self._lib.obj2voxel_set_output_callback(self.handle, cb, None)

    def set_output_memory(self, type_):
        self._lib.obj2voxel_set_output_memory(self.handle, type_.encode())

    def set_output_file(self, path, type_=None):
        p = path.encode()
        self._keep.append(p)
        self._lib.obj2voxel_set_output_file(self.handle, p, None if type_ is None else type_.encode())

    def get_output_memory(self):
        size = C.c_size_t()
        ptr = self._lib.obj2voxel_get_output_memory(self.handle, C.byref(size))
        if not ptr:
            return None
        return bytes(np.ctypeslib.as_array(ptr, shape=(max(size.value, 1),))[:size.value])

    def set_slab(self, z0, z1):
        self._lib.obj2voxel_b200_set_slab(self.handle, z0, z1)

    def set_devices(self, devices):
        """obj2voxel_b200_set_devices: the CUDA devices this job is spread over (one Z-slab each)."""
        arr = (C.c_int32 * max(len(devices), 1))(*[int(d) for d in devices])
        self._lib.obj2voxel_b200_set_devices(self.handle, arr, len(devices))

    # -- run ----------------------------------------------------------------------------------------------------
    def voxelize(self):
        return int(self._lib.obj2voxel_voxelize(self.handle))

    def stats(self):
        s = Stats()
        self._lib.obj2voxel_b200_get_stats(self.handle, C.byref(s))
        return s.as_dict()

    def collected(self):
        """All voxels received by the output callback, sorted ascending by (x, y, z)."""
        if not self.voxels:
            return np.zeros((0, 4), dtype=np.uint32)
        return sort_voxels(np.concatenate(self.voxels, axis=0))

    def free(self):
        if self.handle:
            self._lib.obj2voxel_free(self.handle)
            self.handle = None


def sort_voxels(voxels):
    v = np.asarray(voxels).reshape(-1, 4)
    order = np.lexsort((v[:, 2], v[:, 1], v[:, 0]))
    return v[order]


def make_params(resolution, supersampling=1, strategy=MAX_STRATEGY, bounds=None, unit=None, slab=None, variant=-1,
                prefilter=1, occupancy_path=1, slab_filtered=0, float_records=0, accumulate=0):
    p = Params()
    _lib.load().o2v_b200_default_params(C.byref(p))
    p.resolution = resolution
    p.supersampling = supersampling
    p.strategy = strategy
    if bounds is not None:
        p.bounds_known = 1
        p.bounds = (C.c_float * 6)(*[float(b) for b in bounds])
    if unit is not None:
        p.unit_transform = (C.c_int32 * 9)(*[int(u) for u in unit])
    if slab is not None:
        if int(slab[0]) >= int(slab[1]):
            # (0, 0) is the engine's "whole grid"; an empty slab must be skipped by its rank, never passed down
            raise ValueError("empty slab %r: skip the rank instead (slabs.my_slab returns None for it)" % (slab,))
        p.slab_z0, p.slab_z1 = int(slab[0]), int(slab[1])
    p.variant = variant
    p.prefilter = prefilter
    p.occupancy_path = occupancy_path
    p.slab_filtered = slab_filtered
    p.float_records = float_records
    p.accumulate = accumulate
    return p


def _is_torch_cuda(x):
    return hasattr(x, "is_cuda") and x.is_cuda


class Engine:
    """One GPU's voxelizer (o2v_b200_engine).  Raises DeviceError when no CUDA device is usable."""

    def __init__(self, device=0):
        self._lib = _lib.load()
        self.device = device
        self.handle = self._lib.o2v_b200_engine_create(device)
        if not self.handle:
            raise DeviceError(self._lib.o2v_b200_last_error().decode())
        self.last_stats = None
        self._marshalled = None

    @property
    def sm_count(self):
        return self._lib.o2v_b200_sm_count(self.handle)

    def close(self):
        if self.handle:
            self._lib.o2v_b200_engine_destroy(self.handle)
            self.handle = None

    # -- device-resident path -----------------------------------------------------------------------------------
    def voxelize_device(self, verts, params, uvs=None, types=None, colors=None, texture_ids=None, textures=(),
                        stream=None):
        """verts/uvs/...: torch CUDA tensors (float32 / uint8 / uint32-as-int32, contiguous).  textures: sequence of
        (pixels_cuda_uint8[h,w,c], wrap).  Leaves the result on the device; returns the stats dict."""
        import torch

        def ptr(t, dtype):
            if t is None:
                return None
            assert t.is_cuda and t.is_contiguous() and t.dtype == dtype, "expected a contiguous CUDA tensor"
            return t.data_ptr()

        # the marshalled arguments of the last call are kept: a caller that steps the same tensors again (a benchmark, a
        # simulation loop) pays the ctypes call and nothing else
        key = (verts.data_ptr(), verts.numel(), None if uvs is None else uvs.data_ptr(),
               None if types is None else types.data_ptr(), None if colors is None else colors.data_ptr(),
               None if texture_ids is None else texture_ids.data_ptr(),
               tuple((p.data_ptr(), tuple(p.shape), w) for p, w in textures))
        if self._marshalled is None or self._marshalled[0] != key:
            mesh = Mesh(ptr(verts, torch.float32), ptr(uvs, torch.float32), ptr(types, torch.uint8),
                        ptr(colors, torch.float32), ptr(texture_ids, torch.int32), verts.numel() // 9)
            tex_array = (_lib.Texture * max(len(textures), 1))()
            for i, (pixels, wrap) in enumerate(textures):
                assert pixels.is_cuda and pixels.dtype == torch.uint8 and pixels.is_contiguous()
                h, w, ch = pixels.shape
                tex_array[i] = _lib.Texture(pixels.data_ptr(), w, h, ch, wrap)
            self._marshalled = (key, mesh, tex_array)
        _, mesh, tex_array = self._marshalled
        if stream is None:
            stream = torch.cuda.current_stream(verts.device).cuda_stream
        stats = Stats()
        rc = self._lib.o2v_b200_voxelize_device(self.handle, C.byref(params), C.byref(mesh), tex_array, len(textures),
                                                C.c_void_p(stream), C.byref(stats))
        if rc != 0:
            raise DeviceError("o2v_b200_voxelize_device failed (%d): %s" %
                              (rc, self._lib.o2v_b200_last_error().decode()))
        self.last_stats = stats  # indexable like a dict (stats["voxels"]); .as_dict() for a real one
        return stats

    def result_count(self):
        return int(self._lib.o2v_b200_result_count(self.handle))

    def result_tensor(self):
        """Zero-copy torch view (n, 4) int32 of the engine-owned device result (valid until the next run)."""
        import torch

        n = self.result_count()
        if n == 0:
            return torch.zeros((0, 4), dtype=torch.int32, device="cuda:%d" % self.device)

        class _View:
            pass

        view = _View()
        view.__cuda_array_interface__ = {"shape": (n, 4), "typestr": "<i4",
                                         "data": (int(self._lib.o2v_b200_result_device(self.handle)), False),
                                         "version": 2}
        return torch.as_tensor(view, device="cuda:%d" % self.device)

    def result_floats_tensor(self):
        """Zero-copy torch view (n, 4) float32 (weight, r, g, b) per record of a run with float_records=1, index-aligned
        with result_tensor(); None otherwise."""
        import torch

        ptr = self._lib.o2v_b200_result_floats_device(self.handle)
        n = self.result_count()
        if not ptr or n == 0:
            return None

        class _View:
            pass

        view = _View()
        view.__cuda_array_interface__ = {"shape": (n, 4), "typestr": "<f4", "data": (int(ptr), False), "version": 2}
        return torch.as_tensor(view, device="cuda:%d" % self.device)

    def filter_slab(self, verts, params, stream=None):
        """Multi-GPU ingest (o2v_b200_filter_slab): the triangles of the CUDA tensor `verts` whose z range can reach the
        slab of `params`, as a new (k, 9) CUDA tensor (copied out of the engine-owned array), arbitrary order."""
        import torch

        assert verts.is_cuda and verts.is_contiguous() and verts.dtype == torch.float32
        mesh = Mesh(verts.data_ptr(), None, None, None, None, verts.numel() // 9)
        if stream is None:
            stream = torch.cuda.current_stream(verts.device).cuda_stream
        kept, count = C.c_void_p(), C.c_uint64()
        rc = self._lib.o2v_b200_filter_slab(self.handle, C.byref(params), C.byref(mesh), C.c_void_p(stream),
                                            C.byref(kept), C.byref(count))
        if rc != 0:
            raise DeviceError("o2v_b200_filter_slab failed (%d): %s" % (rc, self._lib.o2v_b200_last_error().decode()))
        n = int(count.value)
        if n == 0:
            return torch.zeros((0, 9), dtype=torch.float32, device=verts.device)

        class _View:
            pass

        view = _View()
        view.__cuda_array_interface__ = {"shape": (n, 9), "typestr": "<f4", "data": (int(kept.value), False),
                                         "version": 2}
        return torch.as_tensor(view, device=verts.device).clone()

    def result_hash(self, stream=None):
        """Order-independent 64-bit checksum of the last result, computed on the device (meshes.record_hash on the host)."""
        h = C.c_uint64()
        rc = self._lib.o2v_b200_result_hash(self.handle, C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise DeviceError(self._lib.o2v_b200_last_error().decode())
        return int(h.value)

    def download(self, out=None, stream=None):
        """Copies the last result to host memory: numpy (n, 4) uint32."""
        n = self.result_count()
        if out is None:
            out = np.empty((n, 4), dtype=np.uint32)
        rc = self._lib.o2v_b200_result_download(self.handle, out.ctypes.data_as(C.c_void_p),
                                                C.c_void_p(stream) if stream else None)
        if rc != 0:
            raise DeviceError(self._lib.o2v_b200_last_error().decode())
        return out[:n]

    # -- host-buffer path ---------------------------------------------------------------------------------------
    def voxelize_host(self, verts, params, uvs=None, types=None, colors=None, texture_ids=None, textures=(),
                      capacity=None):
        """numpy in, numpy out: H2D + kernels + D2H inside the call.  Returns (voxels (n,4) uint32 unsorted, stats)."""
        verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 9)
        keep = [verts]

        def ptr(a, dtype):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dtype)
            keep.append(a)
            return a.ctypes.data_as(C.c_void_p).value

        mesh = Mesh(ptr(verts, np.float32), ptr(uvs, np.float32), ptr(types, np.uint8), ptr(colors, np.float32),
                    ptr(texture_ids, np.uint32), len(verts))
        tex_array = (_lib.Texture * max(len(textures), 1))()
        for i, (pixels, wrap) in enumerate(textures):
            p = np.ascontiguousarray(pixels, dtype=np.uint8)
            keep.append(p)
            h, w, ch = p.shape
            tex_array[i] = _lib.Texture(p.ctypes.data_as(C.c_void_p).value, w, h, ch, wrap)
        stats = Stats()
        count = C.c_uint64()
        if capacity is None:
            capacity = 1 << 16
        while True:
            out = np.empty((capacity, 4), dtype=np.uint32)
            rc = self._lib.o2v_b200_voxelize_host(self.handle, C.byref(params), C.byref(mesh), tex_array,
                                                  len(textures), out.ctypes.data_as(C.POINTER(C.c_uint32)), capacity,
                                                  C.byref(count), C.byref(stats))
            if rc == -5:  # output buffer too small: the exact count is known now
                capacity = int(count.value)
                continue
            if rc != 0:
                raise DeviceError("o2v_b200_voxelize_host failed (%d): %s" %
                                  (rc, self._lib.o2v_b200_last_error().decode()))
            break
        self.last_stats = stats.as_dict()
        return out[:count.value], self.last_stats
