"""Two-device job on a mesh with nine tenths of its triangles below z = 0.2: slab bounds with and without balancing
(O2V_B200_BALANCE=0).  Needs two GPUs."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import obj2voxel_b200 as o2v
from obj2voxel_b200 import meshes
lib = o2v.load(); lib.obj2voxel_set_log_level(o2v._lib.LOG_DEBUG)
low = meshes.random_triangles(1_800_000, 0.002, seed=47); low[:, 2::3] *= np.float32(0.2)
verts = np.concatenate([low, meshes.random_triangles(200_000, 0.002, seed=48)])
import time
for step in range(3):
    import ctypes as C
    counter = o2v._lib.CountingSink(0, 0)
    inst = o2v.Instance(); inst.set_input_triangles(verts)
    lib.obj2voxel_set_output_callback(inst.handle, C.cast(lib.obj2voxel_b200_counting_sink_write, o2v._lib.VOXEL_CALLBACK), C.addressof(counter))
    inst.set_resolution(1024); inst.set_mesh_boundaries(meshes.UNIT_BOUNDS); inst.set_devices([0, 1])
    t0 = time.perf_counter(); err = inst.voxelize(); dt = time.perf_counter() - t0
    print("step", step, "err", err, "voxels", counter.voxels, "ms", dt * 1e3, flush=True)
    inst.free()
