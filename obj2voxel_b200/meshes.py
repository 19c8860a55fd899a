"""Synthetic meshes of BASELINE.json's configs, generated from a counter-based PRNG (splitmix64) so that
numpy (host, tests, CPU baseline) and torch (device, bench) produce bit-identical float32 arrays.

All generators return model-space triangles as float32 (n, 9); random sets live in the unit cube and are meant to be run
with mesh bounds {0,0,0,1,1,1} so the transform is data-independent.
"""
import numpy as np

_GOLDEN = 0x9E3779B97F4A7C15
_M1 = 0xBF58476D1CE4E5B9
_M2 = 0x94D049BB133111EB
_MASK = (1 << 64) - 1

UNIT_BOUNDS = (0.0, 0.0, 0.0, 1.0, 1.0, 1.0)


def _splitmix64_np(index, seed):
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + (index.astype(np.uint64) + np.uint64(1)) * np.uint64(_GOLDEN))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_M1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_M2)
        z = z ^ (z >> np.uint64(31))
    return z


def uniform_np(start, count, seed):
    """float32 U[0,1) values number start .. start+count-1 of stream `seed`: float(x >> 40) * 2^-24."""
    idx = np.arange(start, start + count, dtype=np.uint64)
    return ((_splitmix64_np(idx, seed) >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def _to_i64(x):
    x &= _MASK
    return x - (1 << 64) if x >= (1 << 63) else x


def uniform_torch(start, count, seed, device):
    """Same stream as uniform_np, computed on `device` with wrapping int64 arithmetic."""
    import torch

    def lsr(x, k):  # logical shift right on two's-complement int64
        return (x >> k) & ((1 << (64 - k)) - 1)

    idx = torch.arange(start, start + count, dtype=torch.int64, device=device)
    z = (idx + 1) * _to_i64(_GOLDEN) + _to_i64(seed)
    z = (z ^ lsr(z, 30)) * _to_i64(_M1)
    z = (z ^ lsr(z, 27)) * _to_i64(_M2)
    z = z ^ lsr(z, 31)
    return lsr(z, 40).to(torch.float32) * (2.0 ** -24)


def record_hash(voxels):
    """Order-independent 64-bit checksum of a set of (x, y, z, argb) records (any order, numpy (n, 4) uint32): the sum mod
    2^64 of the splitmix64 finaliser of x + (y << 21) + (z << 42) xor argb * 0x9E3779B97F4A7C15.  The device computes the
    same sum over an engine's result (o2v_b200_result_hash) and ranks add their parts, so a multi-GPU job is checked against
    the reference without gathering a single record."""
    v = np.asarray(voxels).reshape(-1, 4).astype(np.uint64)
    with np.errstate(over="ignore"):
        k = v[:, 0] + (v[:, 1] << np.uint64(21)) + (v[:, 2] << np.uint64(42))
        k = k ^ (v[:, 3] * np.uint64(_GOLDEN))
        k = (k ^ (k >> np.uint64(30))) * np.uint64(_M1)
        k = (k ^ (k >> np.uint64(27))) * np.uint64(_M2)
        k = k ^ (k >> np.uint64(31))
        return int(np.sum(k, dtype=np.uint64))


def single_triangle():
    """cfg1: (0,0,0),(0,0,1),(1,0,0) — reference test/main.cpp:15-19."""
    return np.array([[0, 0, 0, 0, 0, 1, 1, 0, 0]], dtype=np.float32)


def unit_cube():
    """12 triangles from 6 quads exactly as reference test/main.cpp:21-39 + testutil.hpp:88-113 split them."""
    v = np.array([[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1], [1, 0, 0], [1, 0, 1], [1, 1, 0], [1, 1, 1]],
                 dtype=np.float32)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    tris = []
    for q in quads:
        tris.append(np.concatenate([v[q[0]], v[q[1]], v[q[2]]]))
        tris.append(np.concatenate([v[q[2]], v[q[3]], v[q[0]]]))
    return np.array(tris, dtype=np.float32)


def three_planes():
    """Reference test/main.cpp:41-63: x = 0, 0.5, 1 unit squares as quads."""
    tris = []
    for x in (0.0, 0.5, 1.0):
        q = np.array([[x, 0, 0], [x, 0, 1], [x, 1, 1], [x, 1, 0]], dtype=np.float32)
        tris.append(np.concatenate([q[0], q[1], q[2]]))
        tris.append(np.concatenate([q[2], q[3], q[0]]))
    return np.array(tris, dtype=np.float32)


def lumpy_sphere(nlat=188, nlon=187):
    """cfg2: closed UV sphere, radius 0.5 * (1 + 0.1 sin(3 theta) sin(5 phi)); 2 * nlon * (nlat - 1) triangles
    (~70 k for the defaults).  float64 trig rounded once to float32."""
    theta = np.linspace(0.0, np.pi, nlat + 1)
    phi = np.linspace(0.0, 2.0 * np.pi, nlon, endpoint=False)
    t, p = np.meshgrid(theta, phi, indexing="ij")
    r = 0.5 * (1.0 + 0.1 * np.sin(3.0 * t) * np.sin(5.0 * p))
    pts = np.stack([r * np.sin(t) * np.cos(p), r * np.sin(t) * np.sin(p), r * np.cos(t)], axis=-1).astype(np.float32)
    tris = []
    for i in range(nlat):
        a = pts[i]
        b = pts[i + 1]
        a2 = np.roll(a, -1, axis=0)
        b2 = np.roll(b, -1, axis=0)
        if i != 0:
            tris.append(np.concatenate([a, b, a2], axis=1))
        if i != nlat - 1:
            tris.append(np.concatenate([a2, b, b2], axis=1))
    return np.ascontiguousarray(np.concatenate(tris, axis=0), dtype=np.float32)


def _assemble(u, n, extent, xp):
    # u: 12 uniforms per triangle: 3 centre coordinates then 9 vertex offsets in [-extent, extent]
    u = u.reshape(n, 12)
    centre = u[:, 0:3]
    offset = (u[:, 3:12] * 2.0 - 1.0) * extent
    if xp is np:
        return (np.tile(centre, (1, 3)) + offset).astype(np.float32)
    return (centre.repeat(1, 3) + offset).contiguous()


def random_triangles(n, extent, seed=1, start=0):
    """Triangles start .. start+n-1 of the random set: centre U[0,1]^3, vertex offsets U[-extent, extent]^3."""
    u = uniform_np(start * 12, n * 12, seed)
    return _assemble(u, n, np.float32(extent), np)


def random_triangles_torch(n, extent, seed=1, start=0, device="cuda"):
    import torch

    out = torch.empty((n, 9), dtype=torch.float32, device=device)
    step = 1 << 22
    for s in range(0, n, step):
        m = min(step, n - s)
        u = uniform_torch((start + s) * 12, m * 12, seed, device)
        out[s:s + m] = _assemble(u, m, float(np.float32(extent)), torch)
    return out


def random_uvs(n, seed=2, start=0):
    return uniform_np(start * 6, n * 6, seed).reshape(n, 6)


def random_uvs_torch(n, seed=2, start=0, device="cuda"):
    return uniform_torch(start * 6, n * 6, seed, device).reshape(n, 6).contiguous()


def random_texture(width=256, height=256, channels=3, seed=3):
    """PRNG bytes, (height, width, channels) uint8."""
    idx = np.arange(width * height * channels, dtype=np.uint64)
    return (_splitmix64_np(idx, seed) >> np.uint64(56)).astype(np.uint8).reshape(height, width, channels)


def slivers(n, seed=21):
    """Needle triangles (length 0.01 .. 0.6, width 1e-9 .. 1e-5 model units) in the unit cube: their computed normals
    are dominated by rounding, which is where the reference's plane-distance cull (voxelization.cpp:451-458) decides."""
    rng = np.random.default_rng(seed)
    v0 = rng.uniform(0.1, 0.9, (n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    perp = rng.normal(size=(n, 3))
    perp -= (perp * d).sum(1, keepdims=True) * d
    perp /= np.linalg.norm(perp, axis=1, keepdims=True)
    length = rng.uniform(0.01, 0.6, (n, 1))
    width = 10.0 ** rng.uniform(-9, -5, (n, 1))
    v1 = v0 + length * d
    v2 = v0 + rng.uniform(0.2, 0.8, (n, 1)) * length * d + width * perp
    return np.clip(np.concatenate([v0, v1, v2], axis=1), 0.0, 1.0).astype(np.float32)


# BASELINE.json configs made concrete.  `extent` is the vertex offset half-range in model units.
CONFIGS = {
    "cfg1": dict(kind="single", resolution=16, supersampling=1, strategy=0),
    "cfg2": dict(kind="sphere", resolution=256, supersampling=1, strategy=0),
    "cfg3": dict(kind="random", n=1_000_000, extent=0.004, resolution=512, supersampling=1, strategy=1, textured=True),
    "cfg4": dict(kind="random", n=10_000_000, extent=0.001, resolution=1024, supersampling=2, strategy=0),
    "cfg5": dict(kind="random", n=100_000_000, extent=0.25 / 2048, resolution=2048, supersampling=1, strategy=0),
    # the headline workload of bench.py: BASELINE.json's metric is quoted at 1024^3
    "r1024": dict(kind="random", n=10_000_000, extent=0.001, resolution=1024, supersampling=1, strategy=0),
}
