"""Host-to-host probe of BASELINE config 4 through obj2voxel_voxelize() with the DEBUG timing line of the job runner.
Usage: e2e_probe.py [steps] ; environment: O2V_B200_DOWNLOAD, O2V_B200_PIPELINE_PARTS, O2V_B200_HOST_THREADS, O2V_B200_DEVICES"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import obj2voxel_b200 as o2v
from obj2voxel_b200 import meshes
import bench
cfg = bench.workload_spec(os.environ.get("O2V_WORKLOAD", "cfg4"))
dev = torch.device("cuda", 0)
verts = meshes.random_triangles_torch(cfg["n"], cfg["extent"], seed=1, device=dev)
pinned = torch.empty(verts.shape, dtype=verts.dtype, pin_memory=True); pinned.copy_(verts)
del verts; torch.cuda.empty_cache()
host = pinned.numpy()
lib = o2v.load(); lib.obj2voxel_set_log_level(o2v._lib.LOG_DEBUG)
for step in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    inst = o2v.Instance(); inst.set_input_triangles(host)
    import ctypes as C
    counter = o2v._lib.CountingSink(0, 0)
    f = C.cast(lib.obj2voxel_b200_counting_sink_write, o2v._lib.VOXEL_CALLBACK)
    lib.obj2voxel_set_output_callback(inst.handle, f, C.addressof(counter))
    inst.set_resolution(cfg["resolution"]); inst.set_supersampling(cfg["supersampling"]); inst.set_mesh_boundaries(cfg["bounds"])
    t0 = time.perf_counter(); err = inst.voxelize(); dt = time.perf_counter() - t0
    print("step", step, "err", err, "voxels", counter.voxels, "calls", counter.calls, "ms", dt * 1e3, flush=True)
    inst.free()
