"""Small mixed workload for compute-sanitizer (memcheck / racecheck / initcheck): touches every kernel path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import obj2voxel_b200 as o2v
from obj2voxel_b200 import meshes
eng = o2v.Engine(0)
cases = [
    (meshes.unit_cube(), dict(resolution=64), {}),                                             # aligned, big leaves
    (meshes.random_triangles(3000, 0.03), dict(resolution=128, strategy=1, bounds=meshes.UNIT_BOUNDS), {}),
    (meshes.random_triangles(3000, 0.03), dict(resolution=64, supersampling=2, strategy=1, bounds=meshes.UNIT_BOUNDS), {}),
    (meshes.random_triangles(40, 0.4, seed=5), dict(resolution=128, strategy=1), {}),           # deep subdivision
    (meshes.lumpy_sphere(40, 41), dict(resolution=128), {}),                                    # block fold tiles
    (meshes.random_triangles(6000, 0.02, seed=23) * np.float32(0.1), dict(resolution=32, strategy=1, bounds=[0, 0, 0, 1, 1, 1]), {}),  # heavy tiles, long lists
    (meshes.random_triangles(2000, 0.03), dict(resolution=128, strategy=1, bounds=meshes.UNIT_BOUNDS),
     dict(uvs=meshes.random_uvs(2000) * 3 - 1, textures=[(meshes.random_texture(32, 16, 3), 1)])),
    (meshes.random_triangles(5000, 0.002, seed=4), dict(resolution=128, bounds=meshes.UNIT_BOUNDS), {}),  # micro-triangles: thread-per-leaf classifier
    # huge triangles: listed, then walked as (triangle, subtree) items by the huge* kernels (first run: the retry)
    (np.concatenate([meshes.random_triangles(3, 0.45, seed=61), meshes.random_triangles(500, 0.02, seed=62)]),
     dict(resolution=400, strategy=1, bounds=[-0.01, -0.01, -0.01, 1.01, 1.01, 1.01]), {}),
    (np.concatenate([meshes.random_triangles(2, 0.45, seed=63), meshes.random_triangles(300, 0.02, seed=64)]),
     dict(resolution=384, strategy=1, bounds=[-0.01, -0.01, -0.01, 1.01, 1.01, 1.01]),
     dict(uvs=meshes.random_uvs(302, seed=65), textures=[(meshes.random_texture(32, 16, 3), 1)])),
]
for verts, kw, extra in cases:
    for occ in (1, 0):  # all-white meshes: occupancy-only pipeline, then the weighted one; textured: weighted twice
        v, st = eng.voxelize_host(verts, o2v.make_params(occupancy_path=occ, **kw), **extra)
        print(len(v), st["occupancy_path"], st["light_tiles"], st["heavy_tiles"], st["survivors"], flush=True)
# occupancy pipeline without its SAT shortcuts (queue growth + rerun) and on a Z-slab
v, st = eng.voxelize_host(meshes.random_triangles(3000, 0.03), o2v.make_params(resolution=128, prefilter=0, bounds=meshes.UNIT_BOUNDS))
print(len(v), st["occupancy_path"], st["survivors"], flush=True)
v, st = eng.voxelize_host(meshes.random_triangles(3000, 0.03), o2v.make_params(resolution=64, supersampling=2, slab=(40, 104), bounds=meshes.UNIT_BOUNDS))
print(len(v), st["occupancy_path"], st["survivors"], flush=True)
# float records (weighted fold with the debug output), record hash, slab ingest
v, st = eng.voxelize_host(meshes.random_triangles(1500, 0.03), o2v.make_params(resolution=64, strategy=1, float_records=1, bounds=meshes.UNIT_BOUNDS))
print(len(v), st["occupancy_path"], flush=True)
import torch
tv = torch.from_numpy(meshes.random_triangles(4000, 0.02, seed=6)).cuda()
p = o2v.make_params(resolution=128, slab=(64, 128), bounds=meshes.UNIT_BOUNDS)
kept = eng.filter_slab(tv, p)
st = eng.voxelize_device(kept, o2v.make_params(resolution=128, slab=(64, 128), slab_filtered=1, bounds=meshes.UNIT_BOUNDS))
print(kept.shape[0], st["voxels"], hex(eng.result_hash()), flush=True)
# a mesh in pieces (o2v_b200_params::accumulate): every chunk has a bitmap that stays, expand masks with the delivered bits
pieces = np.array_split(meshes.random_triangles(6000, 0.02, seed=9), 3)
for k, piece in enumerate(pieces):
    v, st = eng.voxelize_host(np.ascontiguousarray(piece), o2v.make_params(resolution=64, supersampling=2, accumulate=1 if k == 0 else 2,
                                                                          bounds=meshes.UNIT_BOUNDS), capacity=1 << 20)
    print("piece", k, len(v), flush=True)
eng.close()
# the host-to-host job: parts, packed positions / bitmaps / records over PCIe, staged pageable upload
import os
big = meshes.random_triangles(250_000, 0.004, seed=8)
for mode, parts in (("packed", "3"), ("bitmap", "2"), ("records", "3")):
    os.environ["O2V_B200_DOWNLOAD"] = mode
    os.environ["O2V_B200_PIPELINE_PARTS"] = parts
    inst = o2v.Instance()
    inst.set_input_triangles(big)
    inst.set_output_callback()
    inst.set_resolution(200)
    inst.set_supersampling(2 if mode == "packed" else 1)
    inst.set_mesh_boundaries(meshes.UNIT_BOUNDS)
    err = inst.voxelize()
    print(mode, err, len(inst.collected()), flush=True)
    inst.free()
print("sanitize run done")
