"""Runs one workload device-resident a few times (for ncu / timing experiments).  Usage: profile_run.py [workload] [steps]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import obj2voxel_b200 as o2v
from obj2voxel_b200 import meshes
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = bench.workload_spec(name)
if len(sys.argv) > 3:
    cfg["n"] = int(sys.argv[3])
dev = torch.device("cuda", 0)
eng = o2v.Engine(0)
if cfg["kind"] == "sphere":
    verts = torch.from_numpy(meshes.lumpy_sphere()).to(dev); uvs = None
elif cfg["kind"] == "single":
    verts = torch.from_numpy(meshes.single_triangle()).to(dev); uvs = None
else:
    verts = meshes.random_triangles_torch(cfg["n"], cfg["extent"], seed=1, device=dev)
    uvs = meshes.random_uvs_torch(cfg["n"], seed=2, device=dev) if cfg.get("textured") else None
textures = [(torch.from_numpy(meshes.random_texture(256, 256, 3)).to(dev), o2v.UV_WRAP)] if uvs is not None else []
slab = tuple(int(x) for x in os.environ["O2V_SLAB"].split(",")) if "O2V_SLAB" in os.environ else None
params = o2v.make_params(resolution=cfg["resolution"], supersampling=cfg["supersampling"], strategy=cfg["strategy"],
                         bounds=cfg["bounds"], variant=int(os.environ.get("O2V_VARIANT", "-1")), slab=slab,
                         occupancy_path=int(os.environ.get("O2V_OCC", "1")), prefilter=int(os.environ.get("O2V_PREFILTER", "1")))
if slab is not None and os.environ.get("O2V_PREFILTERED", "0") == "1" and uvs is None:
    # what a rank of a multi-GPU run does: the slab's triangles were distributed once at ingest
    verts = eng.filter_slab(verts, params)
    params = o2v.make_params(resolution=cfg["resolution"], supersampling=cfg["supersampling"], strategy=cfg["strategy"],
                             bounds=cfg["bounds"], slab=slab, slab_filtered=1)
    print("slab triangles:", verts.shape[0], flush=True)
if os.environ.get("O2V_TIMED", "0") != "0":
    for i in range(3):
        eng.voxelize_device(verts, params, uvs=uvs, textures=textures)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = int(os.environ["O2V_TIMED"])
    e0.record()
    for i in range(reps):
        eng.voxelize_device(verts, params, uvs=uvs, textures=textures)
    e1.record()
    torch.cuda.synchronize()
    print("ms per step over %d steps: %.4f" % (reps, e0.elapsed_time(e1) / reps), flush=True)
for i in range(steps):
    st = eng.voxelize_device(verts, params, uvs=uvs, textures=textures)
    torch.cuda.synchronize()
    print(json.dumps({k: st[k] for k in ("voxels", "leaves", "pairs", "light_tiles", "heavy_tiles", "clip_calls",
                                         "contributions", "candidate_voxels", "survivors", "occupancy_path", "ms_total", "ms_setup",
                                         "ms_voxelize", "ms_clip", "ms_classify", "ms_filter", "ms_expand", "undecided_ranges")}), flush=True)
