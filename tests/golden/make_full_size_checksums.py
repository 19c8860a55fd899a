"""Runs the UNMODIFIED reference (oracle/_ref, built from /root/reference) on BASELINE.json's full-size configurations and
records voxel count, CRC32 of the sorted (x, y, z, argb) list and the order-independent 64-bit record hash
(meshes.record_hash: what bench.py all-reduces across ranks) in tests/golden/full_size_checksums.json.  CPU only, takes
minutes to tens of minutes; run in the build container: `python tests/golden/make_full_size_checksums.py cfg2 cfg3 cfg4`.

cfg4 uses supersampling 2: the reference build with the two-line downscale fix (oracle/downscale_fix.sed) is used; the mesh
is MATERIALLESS, so the (unordered) child fold order of that build cannot change any colour.
"""
import json
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload definitions only)
from obj2voxel_b200 import meshes  # noqa: E402
from oracle import refharness as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "full_size_checksums.json")


def main():
    names = sys.argv[1:] or ["cfg2", "cfg3", "cfg4"]
    results = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in names:
        cfg = bench.workload_spec(name)
        verts, uvs = bench.host_mesh(cfg)
        texture = dict(pixels=meshes.random_texture(256, 256, 3), wrap=1) if uvs is not None else None
        t0 = time.time()
        r = R.run_api(verts, cfg["resolution"], uvs=uvs, texture=texture, supersampling=cfg["supersampling"],
                      strategy=cfg["strategy"], bounds=cfg["bounds"], workers=R.hardware_threads(), collect=True,
                      patched_downscale=cfg["supersampling"] == 2)
        v = np.ascontiguousarray(r["voxels"])
        results[name] = dict(triangles=int(len(verts)), resolution=cfg["resolution"],
                             supersampling=cfg["supersampling"], strategy=cfg["strategy"], voxels=int(len(v)),
                             crc32=int(zlib.crc32(v.tobytes())), hash64=meshes.record_hash(v), reference_seconds=round(time.time() - t0, 1),
                             reference="patched downscale" if cfg["supersampling"] == 2 else "unmodified")
        print(name, results[name], flush=True)
        json.dump(results, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
