import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import obj2voxel_b200 as o2v
from obj2voxel_b200 import meshes
import bench
cfg = bench.workload_spec("cfg4")
dev = torch.device("cuda", 0)
verts = meshes.random_triangles_torch(cfg["n"], cfg["extent"], seed=1, device=dev)
pinned = torch.empty(verts.shape, dtype=verts.dtype, pin_memory=True); pinned.copy_(verts)
host = pinned.numpy()
lib = o2v.load(); lib.obj2voxel_set_log_level(o2v._lib.LOG_DEBUG)
for step in range(4):
    inst = o2v.Instance(); inst.set_input_triangles(host)
    n = [0]
    def cb(_d, _q, c, n=n):
        n[0] += c; return True
    f = o2v._lib.VOXEL_CALLBACK(cb); lib.obj2voxel_set_output_callback(inst.handle, f, None)
    inst.set_resolution(1024); inst.set_supersampling(2); inst.set_mesh_boundaries(cfg["bounds"])
    t0 = time.perf_counter(); err = inst.voxelize(); dt = time.perf_counter() - t0
    print("step", step, "err", err, "voxels", n[0], "ms", dt * 1e3, flush=True)
    inst.free()
