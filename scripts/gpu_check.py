"""Ad-hoc GPU parity sweep with diagnostics (development aid; the graded checks live in tests/)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import obj2voxel_b200 as o2v
from obj2voxel_b200 import meshes
from oracle import oracle

def diff(name, got, want, stats):
    got = o2v.sort_voxels(got)
    ok = got.shape == want.shape and np.array_equal(got, want)
    print("%-34s %s gpu=%d oracle=%d leaves=%d pairs=%d tiles=%d(light %d) clips=%d contrib=%d ms=%.3f (setup %.3f vox %.3f)" % (
        name, "OK " if ok else "BAD", len(got), len(want), stats["leaves"], stats["pairs"], stats["active_tiles"], stats["light_tiles"],
        stats["clip_calls"], stats["contributions"], stats["ms_total"], stats["ms_setup"], stats["ms_voxelize"]), flush=True)
    if not ok:
        gs = set(map(tuple, got[:, :3].tolist())); ws = set(map(tuple, want[:, :3].tolist()))
        print("   missing %d extra %d" % (len(ws - gs), len(gs - ws)))
        print("   missing sample", sorted(ws - gs)[:5], "extra sample", sorted(gs - ws)[:5])
        if gs == ws:
            d = np.nonzero(got[:, 3] != want[:, 3])[0]
            print("   colour mismatches", len(d), [(got[i].tolist(), hex(want[i, 3])) for i in d[:5]])
    return ok

def main():
    eng = o2v.Engine(0)
    rng = np.random.default_rng(7)
    allok = True
    cases = []
    cases.append(("cfg1 single r16", meshes.single_triangle(), dict(resolution=16), {}))
    cases.append(("cube r64", meshes.unit_cube(), dict(resolution=64), {}))
    cases.append(("cube r128", meshes.unit_cube(), dict(resolution=128), {}))
    cases.append(("planes r32", meshes.three_planes(), dict(resolution=32), {}))
    v = meshes.random_triangles(3000, 0.03)
    cases.append(("rand3k r128 max", v, dict(resolution=128, strategy=0, bounds=meshes.UNIT_BOUNDS), {}))
    cases.append(("rand3k r128 blend", v, dict(resolution=128, strategy=1, bounds=meshes.UNIT_BOUNDS), {}))
    cases.append(("rand3k r100 blend nobounds", v, dict(resolution=100, strategy=1), {}))
    big = meshes.random_triangles(60, 0.4, seed=5)
    cases.append(("big60 r256 blend", big, dict(resolution=256, strategy=1), {}))
    uv = meshes.random_uvs(3000) * 3 - 1
    tex = meshes.random_texture(64, 48, 3)
    cases.append(("tex rgb blend", v, dict(resolution=128, strategy=1, bounds=meshes.UNIT_BOUNDS), dict(uvs=uv, tex=(tex, 1))))
    tex4 = meshes.random_texture(33, 17, 4, seed=9)
    cases.append(("tex argb clamp max", v, dict(resolution=64, strategy=0, bounds=meshes.UNIT_BOUNDS), dict(uvs=uv, tex=(tex4, 0))))
    types = rng.integers(1, 3, len(v)).astype(np.uint8); cols = rng.random((len(v), 3)).astype(np.float32)
    cases.append(("untextured colours blend", v, dict(resolution=128, strategy=1, bounds=meshes.UNIT_BOUNDS), dict(types=types, colors=cols)))
    cases.append(("ss2 max", v, dict(resolution=64, supersampling=2, strategy=0, bounds=meshes.UNIT_BOUNDS), {}))
    cases.append(("ss2 blend", v, dict(resolution=64, supersampling=2, strategy=1, bounds=meshes.UNIT_BOUNDS), {}))
    sph = meshes.lumpy_sphere(60, 61)
    cases.append(("sphere r128", sph, dict(resolution=128), {}))
    for name, verts, kw, extra in cases:
        for prefilter in (1, 0):
            params = o2v.make_params(prefilter=prefilter, **kw)
            textures = [extra["tex"]] if "tex" in extra else []
            t0 = time.time()
            got, stats = eng.voxelize_host(verts, params, uvs=extra.get("uvs"), types=extra.get("types"),
                                           colors=extra.get("colors"), textures=textures)
            okw = dict(strategy=kw.get("strategy", 0), bounds=kw.get("bounds"), supersampling=kw.get("supersampling", 1))
            if "tex" in extra:
                okw.update(uvs=extra["uvs"], texture=dict(pixels=extra["tex"][0], wrap=extra["tex"][1]))
            if "types" in extra:
                okw.update(types=extra["types"], colors=extra["colors"])
            want = oracle.voxelize(verts, kw["resolution"], **okw)["voxels"]
            allok &= diff(name + (" pf" if prefilter else " nopf"), got, want, stats)
    # slab check: union of two slabs == whole
    params = o2v.make_params(resolution=128, strategy=1, bounds=meshes.UNIT_BOUNDS)
    whole, _ = eng.voxelize_host(v, params)
    parts = []
    for slab in ((0, 64), (64, 128)):
        p = o2v.make_params(resolution=128, strategy=1, bounds=meshes.UNIT_BOUNDS, slab=slab)
        g, _ = eng.voxelize_host(v, p); parts.append(g)
    u = o2v.sort_voxels(np.concatenate(parts)); w = o2v.sort_voxels(whole)
    ok = u.shape == w.shape and np.array_equal(u, w); allok &= ok
    print("slab union == whole:", ok)
    print("ALL OK" if allok else "SOME FAILED")
    return 0 if allok else 1

if __name__ == "__main__":
    sys.exit(main())
