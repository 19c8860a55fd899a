"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libo2vref.so, built by oracle/Makefile from
/root/reference).  Run in the build container only: `python tests/golden/make_golden.py`.  The fixtures travel to the GPU
box; the reference does not.

Each fixture stores the inputs and, from the reference: the sorted (x,y,z,argb) voxel list of the public C API and the
float (weight,r,g,b) WeightedColor values + mesh transform of the internal pipeline (bit patterns as uint32).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from obj2voxel_b200 import meshes  # noqa: E402  (generators only; no product compute is involved)
from oracle import refharness as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def save(name, verts, resolution, uvs=None, texture=None, types=None, colors=None, **kw):
    api_kw = {k: v for k, v in kw.items() if k in ("supersampling", "strategy", "bounds", "unit")}
    data = dict(verts=np.asarray(verts, np.float32), resolution=np.int64(resolution),
                strategy=np.int64(kw.get("strategy", 0)), supersampling=np.int64(kw.get("supersampling", 1)))
    if kw.get("bounds") is not None:
        data["bounds"] = np.asarray(kw["bounds"], np.float32)
    if kw.get("unit") is not None:
        data["unit"] = np.asarray(kw["unit"], np.int32)
    if uvs is not None:
        data["uvs"] = np.asarray(uvs, np.float32)
        data["tex_pixels"] = texture["pixels"]
        data["tex_wrap"] = np.int64(texture["wrap"])
    if types is not None:
        data["types"] = np.asarray(types, np.uint8)
        data["colors"] = np.asarray(colors, np.float32)
    patched = kw.get("patched_downscale", False)
    if types is None:
        # the public API cannot express UNTEXTURED triangles (set_triangle_colored is MATERIALLESS)
        api = R.run_api(verts, resolution, uvs=uvs, texture=texture, patched_downscale=patched, **api_kw)
        data["api_voxels"] = api["voxels"]
        data["api_sink_calls"] = np.int64(api["sink_calls"])
    internal = R.run_internal(verts, resolution, uvs=uvs, texture=texture, types=types, colors=colors,
                              apply_downscale=patched, patched_downscale=patched, **api_kw)
    data["int_xyz"] = internal["xyz"]
    data["int_wrgb_bits"] = internal["wrgb"].view(np.uint32)
    data["transform_bits"] = internal["transform"].view(np.uint32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
    print("%-28s %8d voxels" % (name, len(internal["xyz"])))


def main(only=None):
    global save
    if only:
        _save = save
        save = lambda name, *a, **k: _save(name, *a, **k) if name in only else None  # noqa: E731
    rng = np.random.default_rng(2024)
    save("cfg1_single_r16_max", meshes.single_triangle(), 16, strategy=0)
    save("cfg1_single_r16_blend", meshes.single_triangle(), 16, strategy=1)
    save("cube_r16", meshes.unit_cube(), 16)
    save("planes_r32", meshes.three_planes(), 32)
    v = meshes.random_triangles(500, 0.04, seed=11)
    save("rand500_r64_max", v, 64, strategy=0, bounds=meshes.UNIT_BOUNDS)
    save("rand500_r64_blend", v, 64, strategy=1, bounds=meshes.UNIT_BOUNDS)
    save("rand500_r96_blend_autobounds_perm", v, 96, strategy=1, unit=[0, 0, 1, 1, 0, 0, 0, -1, 0])
    uv = meshes.random_uvs(500, seed=12) * 3 - 1
    save("rand500_r64_tex_rgb_wrap_blend", v, 64, uvs=uv, texture=dict(pixels=meshes.random_texture(32, 24, 3), wrap=1),
         strategy=1, bounds=meshes.UNIT_BOUNDS)
    save("rand500_r64_tex_argb_clamp_max", v, 64, uvs=uv,
         texture=dict(pixels=meshes.random_texture(17, 9, 4, seed=5), wrap=0), strategy=0, bounds=meshes.UNIT_BOUNDS)
    types = rng.integers(1, 3, len(v)).astype(np.uint8)
    cols = rng.random((len(v), 3)).astype(np.float32)
    save("rand500_r64_untextured_blend", v, 64, types=types, colors=cols, strategy=1, bounds=meshes.UNIT_BOUNDS)
    save("big12_r128_blend_subdivided", meshes.random_triangles(12, 0.45, seed=13), 128, strategy=1)
    save("sphere24_r64_max", meshes.lumpy_sphere(24, 25), 64, strategy=0)
    # needle triangles: the plane-distance cull decides occupancy (their normals are rounding noise)
    save("slivers300_r256_blend", meshes.slivers(300), 256, strategy=1, bounds=meshes.UNIT_BOUNDS)
    save("slivers300_r64_max_autobounds", meshes.slivers(300, seed=22), 64, strategy=0)
    # supersampling: the unmodified reference yields zero voxels; pinned two ways
    save("rand500_r64_presample_of_ss2", v, 64, strategy=0, bounds=meshes.UNIT_BOUNDS)  # == resolution 32, ss 2 pre-downscale
    save("rand500_r32_ss2_patched_max", v, 32, supersampling=2, strategy=0, bounds=meshes.UNIT_BOUNDS,
         patched_downscale=True)
    save("rand500_r32_ss2_patched_blend", v, 32, supersampling=2, strategy=1, bounds=meshes.UNIT_BOUNDS,
         patched_downscale=True)


if __name__ == "__main__":
    main(set(sys.argv[1:]))  # optional: names of the fixtures to (re)generate
