"""A low-poly model at a high resolution: a cube rotated off the axes (12 triangles, each ~0.7 of the grid across) and the
same cube plus 200 k small triangles.  Prints the device-resident step with the library in use; run it once with the
default build and once with an A/B build that never treats a triangle as huge
(scripts/build_variants.sh nohuge "-DO2V_HUGE_ROOT_VOLUME=(1ull<<62)"; O2V_B200_LIB=.variants/libo2v_nohuge.so)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import obj2voxel_b200 as o2v
from obj2voxel_b200 import meshes

def rotated_cube():
    c = meshes.unit_cube().reshape(-1, 3).astype(np.float64) - 0.5
    a, b = 0.6, 0.35
    rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    rx = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
    return (c @ (rz @ rx).T * 0.55 + 0.5).astype(np.float32).reshape(-1, 9)

dev = torch.device("cuda", 0)
eng = o2v.Engine(0)
cube = rotated_cube()
for name, verts in (("cube", cube), ("cube + 200k small", np.concatenate([cube, meshes.random_triangles(200000, 0.004, seed=5)]))):
    for res in (1024, 2048):
        for occ in (1, 0):
            v = torch.from_numpy(verts).to(dev)
            p = o2v.make_params(resolution=res, strategy=o2v.MAX_STRATEGY, bounds=(0, 0, 0, 1, 1, 1), occupancy_path=occ)
            for _ in range(2):
                st = eng.voxelize_device(v, p)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                st = eng.voxelize_device(v, p)
            e1.record(); torch.cuda.synchronize()
            print(json.dumps({"mesh": name, "resolution": res, "pipeline": "occupancy" if occ else "weighted",
                              "ms_per_step": round(e0.elapsed_time(e1) / 3, 3), "leaves": st["leaves"], "voxels": st["voxels"],
                              "hash": eng.result_hash()}), flush=True)
