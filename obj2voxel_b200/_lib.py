"""ctypes declarations for libobj2voxel_b200.so (include/obj2voxel.h + include/obj2voxel_b200.h).

The library is the product: hand-written sm_100a kernels behind a C ABI.  This module only loads it — it never
substitutes another implementation.  A missing library raises; a missing GPU makes engine creation / voxelization fail
loudly (OBJ2VOXEL_ERR_DEVICE, o2v_b200_last_error()).
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# O2V_B200_LIB: another build of the same library (kernel A/B experiments, scripts/build_variants.sh)
LIB_PATH = os.environ.get("O2V_B200_LIB") or os.path.join(_HERE, "libobj2voxel_b200.so")
_LIB = None

# include/obj2voxel.h enum constants
MAX_STRATEGY, BLEND_STRATEGY = 0, 1
UV_CLAMP, UV_WRAP = 0, 1
LOG_SILENT, LOG_ERROR, LOG_WARNING, LOG_INFO, LOG_DEBUG = 0, 1, 2, 3, 4
ERR_OK, ERR_NO_INPUT, ERR_NO_OUTPUT, ERR_NO_RESOLUTION = 0, 1, 2, 3
ERR_IO_OPEN_INPUT, ERR_IO_OPEN_OUTPUT, ERR_IO_WRITE, ERR_DOUBLE_VOXELIZATION, ERR_DEVICE = 4, 5, 6, 7, 8

TRIANGLE_CALLBACK = C.CFUNCTYPE(C.c_bool, C.c_void_p, C.c_void_p)
VOXEL_CALLBACK = C.CFUNCTYPE(C.c_bool, C.c_void_p, C.POINTER(C.c_uint32), C.c_size_t)
LOG_CALLBACK = C.CFUNCTYPE(C.c_bool, C.c_void_p, C.c_char_p, C.c_ubyte)


class Params(C.Structure):
    _fields_ = [("resolution", C.c_uint32), ("supersampling", C.c_uint32), ("strategy", C.c_uint32),
                ("bounds_known", C.c_uint32), ("bounds", C.c_float * 6), ("unit_transform", C.c_int32 * 9),
                ("slab_z0", C.c_uint32), ("slab_z1", C.c_uint32), ("variant", C.c_int32), ("prefilter", C.c_int32),
                ("float_records", C.c_int32), ("slab_filtered", C.c_int32), ("occupancy_path", C.c_int32),
                ("accumulate", C.c_int32)]


class Mesh(C.Structure):
    _fields_ = [("verts", C.c_void_p), ("uvs", C.c_void_p), ("types", C.c_void_p), ("colors", C.c_void_p),
                ("texture_ids", C.c_void_p), ("count", C.c_uint64)]


class Texture(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("channels", C.c_uint32),
                ("wrap", C.c_uint32)]


class ArraySource(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("count", C.c_size_t), ("next", C.c_size_t)]


class CountingSink(C.Structure):
    _fields_ = [("voxels", C.c_uint64), ("calls", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [("voxels", C.c_uint64), ("leaves", C.c_uint64), ("pairs", C.c_uint64), ("active_tiles", C.c_uint64),
                ("candidate_voxels", C.c_uint64), ("clip_calls", C.c_uint64), ("contributions", C.c_uint64),
                ("dropped_triangles", C.c_uint64), ("depth_overflow", C.c_uint64), ("out_capacity", C.c_uint64),
                ("ms_total", C.c_float), ("ms_setup", C.c_float), ("ms_voxelize", C.c_float),
                ("transform", C.c_float * 12), ("kernel_launches", C.c_int32), ("voxelize_launches", C.c_int32),
                ("light_tiles", C.c_uint64), ("heavy_tiles", C.c_uint64), ("survivors", C.c_uint64),
                ("ms_clip", C.c_float), ("occupancy_path", C.c_int32),
                ("ms_classify", C.c_float), ("reserved", C.c_float), ("slab_triangles", C.c_uint64),
                ("undecided_ranges", C.c_uint64), ("ms_filter", C.c_float), ("ms_expand", C.c_float),
                ("download_bytes", C.c_uint64)]

    def as_dict(self):
        d = {name: getattr(self, name) for name, _ in self._fields_ if name != "transform"}
        d["transform"] = list(self.transform)
        return d

    def __getitem__(self, name):  # stats["voxels"]: the struct itself serves as the mapping (no dict per run)
        if name == "transform":
            return list(self.transform)
        try:
            return getattr(self, name)
        except AttributeError:
            raise KeyError(name)

    def keys(self):
        return [name for name, _ in self._fields_]


# every symbol include/obj2voxel.h declares (35) ...
REFERENCE_SYMBOLS = [
    "obj2voxel_alloc", "obj2voxel_free", "obj2voxel_set_log_level", "obj2voxel_set_log_callback",
    "obj2voxel_get_log_level", "obj2voxel_set_resolution", "obj2voxel_set_supersampling",
    "obj2voxel_set_color_strategy", "obj2voxel_set_texture", "obj2voxel_set_input_file",
    "obj2voxel_set_input_callback", "obj2voxel_set_output_file", "obj2voxel_set_output_memory",
    "obj2voxel_set_output_callback", "obj2voxel_set_parallel", "obj2voxel_set_unit_transform",
    "obj2voxel_set_mesh_boundaries", "obj2voxel_get_resolution", "obj2voxel_get_chunk_size",
    "obj2voxel_get_output_memory", "obj2voxel_set_triangle_basic", "obj2voxel_set_triangle_colored",
    "obj2voxel_set_triangle_textured", "obj2voxel_texture_alloc", "obj2voxel_texture_free",
    "obj2voxel_texture_load_from_file", "obj2voxel_texture_load_from_memory", "obj2voxel_texture_load_pixels",
    "obj2voxel_teture_set_uv_mode", "obj2voxel_texture_get_meta", "obj2voxel_texture_get_pixels",
    "obj2voxel_run_worker", "obj2voxel_stop_workers", "obj2voxel_get_worker_count", "obj2voxel_voxelize",
]
# ... and include/obj2voxel_b200.h
ADDITIVE_SYMBOLS = [
    "o2v_b200_engine_create", "o2v_b200_engine_destroy", "o2v_b200_last_error", "o2v_b200_sm_count",
    "o2v_b200_default_params", "o2v_b200_voxelize_device", "o2v_b200_result_device", "o2v_b200_result_count",
    "o2v_b200_result_download", "o2v_b200_result_floats_device", "o2v_b200_voxelize_host", "obj2voxel_b200_set_input_triangles",
    "obj2voxel_b200_set_slab", "obj2voxel_b200_get_stats", "o2v_b200_plan_parts", "o2v_b200_result_hash", "o2v_b200_filter_slab", "obj2voxel_b200_set_devices", "o2v_b200_expand_bitmaps", "o2v_b200_expand_packed", "o2v_b200_scan_chunk_bitmap", "obj2voxel_b200_array_source_next", "obj2voxel_b200_counting_sink_write",
    "o2v_b200_plan_slabs",
]


def build():
    """Compile the CUDA extension in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(_HERE, "csrc")])


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libobj2voxel_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(obj2voxel_b200 has no fallback implementation)")
    lib = C.CDLL(LIB_PATH)
    vp, u32, sz = C.c_void_p, C.c_uint32, C.c_size_t
    fp = C.POINTER(C.c_float)
    sig = {
        "obj2voxel_alloc": (vp, []),
        "obj2voxel_free": (None, [vp]),
        "obj2voxel_set_log_level": (None, [C.c_ubyte]),
        "obj2voxel_set_log_callback": (None, [vp, vp]),
        "obj2voxel_get_log_level": (C.c_ubyte, []),
        "obj2voxel_set_resolution": (None, [vp, u32]),
        "obj2voxel_set_supersampling": (None, [vp, u32]),
        "obj2voxel_set_color_strategy": (None, [vp, C.c_ubyte]),
        "obj2voxel_set_texture": (None, [vp, vp]),
        "obj2voxel_set_input_file": (None, [vp, C.c_char_p, C.c_char_p]),
        "obj2voxel_set_input_callback": (None, [vp, TRIANGLE_CALLBACK, vp]),
        "obj2voxel_set_output_file": (None, [vp, C.c_char_p, C.c_char_p]),
        "obj2voxel_set_output_memory": (None, [vp, C.c_char_p]),
        "obj2voxel_set_output_callback": (None, [vp, VOXEL_CALLBACK, vp]),
        "obj2voxel_set_parallel": (None, [vp, C.c_bool]),
        "obj2voxel_set_unit_transform": (None, [vp, C.POINTER(C.c_int)]),
        "obj2voxel_set_mesh_boundaries": (None, [vp, fp]),
        "obj2voxel_get_resolution": (u32, [vp]),
        "obj2voxel_get_chunk_size": (u32, [vp]),
        "obj2voxel_get_output_memory": (C.POINTER(C.c_ubyte), [vp, C.POINTER(sz)]),
        "obj2voxel_set_triangle_basic": (None, [vp, fp]),
        "obj2voxel_set_triangle_colored": (None, [vp, fp, fp]),
        "obj2voxel_set_triangle_textured": (None, [vp, fp, fp, vp]),
        "obj2voxel_texture_alloc": (vp, []),
        "obj2voxel_texture_free": (None, [vp]),
        "obj2voxel_texture_load_from_file": (C.c_bool, [vp, C.c_char_p, C.c_char_p]),
        "obj2voxel_texture_load_from_memory": (C.c_bool, [vp, C.c_char_p, sz, C.c_char_p]),
        "obj2voxel_texture_load_pixels": (C.c_bool, [vp, C.c_char_p, sz, sz, sz]),
        "obj2voxel_teture_set_uv_mode": (None, [vp, C.c_ubyte]),
        "obj2voxel_texture_get_meta": (None, [vp, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz)]),
        "obj2voxel_texture_get_pixels": (None, [vp, C.c_char_p]),
        "obj2voxel_run_worker": (None, [vp]),
        "obj2voxel_stop_workers": (None, [vp]),
        "obj2voxel_get_worker_count": (u32, [vp]),
        "obj2voxel_voxelize": (C.c_ubyte, [vp]),
        "o2v_b200_engine_create": (vp, [C.c_int]),
        "o2v_b200_engine_destroy": (None, [vp]),
        "o2v_b200_last_error": (C.c_char_p, []),
        "o2v_b200_sm_count": (C.c_int, [vp]),
        "o2v_b200_default_params": (None, [C.POINTER(Params)]),
        "o2v_b200_voxelize_device": (C.c_int, [vp, C.POINTER(Params), C.POINTER(Mesh), C.POINTER(Texture), u32, vp,
                                               C.POINTER(Stats)]),
        "o2v_b200_result_device": (vp, [vp]),
        "o2v_b200_result_count": (C.c_uint64, [vp]),
        "o2v_b200_result_download": (C.c_int, [vp, vp, vp]),
        "o2v_b200_result_floats_device": (vp, [vp]),
        "o2v_b200_result_hash": (C.c_int, [vp, vp, C.POINTER(C.c_uint64)]),
        "o2v_b200_filter_slab": (C.c_int, [vp, C.POINTER(Params), C.POINTER(Mesh), vp, C.POINTER(vp),
                                           C.POINTER(C.c_uint64)]),
        "o2v_b200_voxelize_host": (C.c_int, [vp, C.POINTER(Params), C.POINTER(Mesh), C.POINTER(Texture), u32,
                                             C.POINTER(C.c_uint32), C.c_uint64, C.POINTER(C.c_uint64),
                                             C.POINTER(Stats)]),
        "obj2voxel_b200_set_input_triangles": (None, [vp, fp, fp, sz, vp]),
        "obj2voxel_b200_set_slab": (None, [vp, u32, u32]),
        "obj2voxel_b200_set_devices": (None, [vp, C.POINTER(C.c_int32), u32]),
        "o2v_b200_expand_bitmaps": (C.c_uint64, [vp, vp, vp, u32, u32, u32, vp]),
        "o2v_b200_expand_packed": (None, [vp, C.c_int32, C.c_uint64, vp]),
        "o2v_b200_scan_chunk_bitmap": (C.c_uint64, [vp, u32, u32, u32, u32, vp, C.c_uint64]),
        "obj2voxel_b200_array_source_next": (C.c_bool, [vp, vp]),
        "obj2voxel_b200_counting_sink_write": (C.c_bool, [vp, vp, sz]),
        "obj2voxel_b200_get_stats": (None, [vp, C.POINTER(Stats)]),
        "o2v_b200_plan_parts": (u32, [u32, u32, u32, C.c_uint64, C.c_int32, C.POINTER(C.c_uint32), u32]),
        "o2v_b200_plan_slabs": (None, [u32, u32, u32, u32, u32, C.POINTER(C.c_uint64), u32, C.POINTER(C.c_uint32)]),
    }
    for name, (restype, argtypes) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _LIB = lib
    return lib
