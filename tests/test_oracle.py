"""CPU: pins the oracle (oracle/o2v_oracle.c) against the reference — through the committed golden fixtures generated
from the unmodified reference build, through the reference's own test expectations (test/main.cpp), through voxelio's
known-answer tests for the Morton layout, and (when oracle/_ref exists, i.e. in the build container) live."""
import numpy as np
import pytest

from conftest import golden_names, load_golden, oracle_kwargs
from obj2voxel_b200 import meshes
from oracle import oracle, refharness

PATCHED = [n for n in golden_names() if "patched" in n]
EXACT = [n for n in golden_names() if "patched" not in n]


@pytest.mark.parametrize("name", EXACT)
def test_oracle_matches_golden_bit_exact(name):
    g = load_golden(name)
    r = oracle.voxelize(g["verts"], int(g["resolution"]), **oracle_kwargs(g))
    assert np.array_equal(r["transform"].view(np.uint32), g["transform_bits"])
    assert np.array_equal(r["xyz"], g["int_xyz"])  # occupancy + indices
    assert np.array_equal(r["wrgb"].view(np.uint32), g["int_wrgb_bits"])  # float weight and RGB, 0 ulp
    if "api_voxels" in g:
        assert np.array_equal(r["voxels"], g["api_voxels"])  # ARGB8 as the public API emits it


@pytest.mark.parametrize("name", PATCHED)
def test_oracle_supersampling_against_patched_reference(name):
    """Downscale follows the INTENDED semantics; the patched reference folds children in unordered_map
    order, so occupancy must match exactly, MAX weights exactly (max is order-free), BLEND colours to 1 LSB of ARGB8."""
    g = load_golden(name)
    r = oracle.voxelize(g["verts"], int(g["resolution"]), **oracle_kwargs(g))
    assert np.array_equal(r["xyz"], g["int_xyz"])
    ref = g["int_wrgb_bits"].view(np.float32)
    if int(g["strategy"]) == 0:
        assert np.array_equal(r["wrgb"][:, 0], ref[:, 0])
    else:
        assert np.allclose(r["wrgb"], ref, rtol=2e-6, atol=0)
    got = r["voxels"][:, 3].astype(np.int64)
    want = g["api_voxels"][:, 3].astype(np.int64)
    for shift in (0, 8, 16):
        assert np.max(np.abs(((got >> shift) & 255) - ((want >> shift) & 255))) <= (0 if int(g["strategy"]) == 0 else 1)


def test_presample_equals_double_resolution():
    """the pre-downscale set at (R, ss=2) is what the reference computes at resolution 2R, ss=1."""
    g = load_golden("rand500_r64_presample_of_ss2")
    r = oracle.voxelize(g["verts"], 32, supersampling=2, downscale=False, strategy=0, bounds=g["bounds"].tolist())
    assert np.array_equal(r["xyz"], g["int_xyz"])
    assert np.array_equal(r["wrgb"].view(np.uint32), g["int_wrgb_bits"])


def expected_unit_cube_voxels(r):  # reference test/main.cpp:120-126
    return 8 + 12 * (r - 2) + 6 * (r - 2) * (r - 2)


@pytest.mark.parametrize("resolution", [16, 64, 128])
def test_reference_unit_cube_counts(resolution):  # reference test/main.cpp:128-156,194-208
    r = oracle.voxelize(meshes.unit_cube(), resolution)
    assert len(r["xyz"]) == expected_unit_cube_voxels(resolution)


@pytest.mark.parametrize("resolution", [32, 128])
def test_reference_three_planes_counts(resolution):  # reference test/main.cpp:225-252
    r = oracle.voxelize(meshes.three_planes(), resolution)
    assert len(r["xyz"]) == 3 * resolution * resolution


def test_cfg1_golden_set():  # SURVEY §8c golden: {(x,0,z): x+z <= 15}, all white
    r = oracle.voxelize(meshes.single_triangle(), 16)
    want = sorted((x, 0, z) for x in range(16) for z in range(16) if x + z <= 15)
    assert [tuple(p) for p in r["xyz"].tolist()] == want
    assert np.all(r["argb"] == 0xFFFFFFFF)
    assert r["contributions"] == 136


def test_weights_are_multiples_of_whole_triangle_area():
    tri = np.array([[1.3, 2.1, 0.7, 9.2, 3.3, 4.4, 2.2, 8.8, 6.1]], dtype=np.float32)
    r = oracle.voxelize(tri, 16, bounds=[0, 0, 0, 16.5 / 15.5 * 16, 16.5 / 15.5 * 16, 16.5 / 15.5 * 16])
    w = np.sort(np.unique(r["wrgb"][:, 0]))
    ratios = w / w[0]
    assert np.allclose(ratios, np.round(ratios), rtol=1e-5)


def test_morton_layout_known_answers():  # voxelio/test/test_bits.cpp:316-325: x is the MSB of each triple
    assert oracle.ileave3(1, 0, 0) == 4 and oracle.ileave3(0, 1, 0) == 2 and oracle.ileave3(0, 0, 1) == 1
    assert oracle.ileave3(0b1111, 0, 0) == 0b100100100100
    assert oracle.ileave3(0, 0b1111, 0b1111) == 0b011011011011
    rng = np.random.default_rng(3)
    for x, y, z in rng.integers(0, 1 << 21, (200, 3)).tolist():
        assert oracle.dileave3(oracle.ileave3(x, y, z)) == (x, y, z)  # test_bits.cpp:402-425 round trip


def test_argb_quantisation_truncates():  # voxelio color.hpp:165-173: u8(clamp01(x) * 255), not rounding
    assert oracle.quantize_argb([1.0, 0.5, 0.0]) == 0xFFFF7F00
    assert oracle.quantize_argb([2.0, -1.0, 0.999]) == 0xFFFF00FE


def test_empty_and_degenerate_inputs():
    assert len(oracle.voxelize(np.zeros((0, 9), np.float32), 16)["xyz"]) == 0
    # zero-area triangles carry weight 0 and never reach the voxel map (voxelization.cpp:466)
    flat = np.array([[0, 0, 0, 1, 1, 1, 2, 2, 2], [0, 0, 0, 0, 0, 0, 1, 0, 0]], dtype=np.float32)
    assert len(oracle.voxelize(flat, 16)["xyz"]) == 0


def test_max_ties_keep_lowest_triangle_index():
    tri = meshes.single_triangle()
    both = np.concatenate([tri, tri])
    types = np.array([2, 2], dtype=np.uint8)
    colors = np.array([[1, 0, 0], [0, 1, 0]], dtype=np.float32)
    r = oracle.voxelize(both, 16, types=types, colors=colors, strategy=0)
    assert np.all(r["argb"] == 0xFFFF0000)


@pytest.mark.skipif(not refharness.available(), reason="oracle/_ref is only built where /root/reference exists")
def test_oracle_matches_live_reference_fuzz():
    rng = np.random.default_rng(99)
    for case in range(6):
        n = int(rng.integers(50, 400))
        v = meshes.random_triangles(n, float(rng.choice([0.01, 0.05, 0.2])), seed=100 + case)
        res = int(rng.choice([32, 64, 128]))
        strategy = case % 2
        uv = meshes.random_uvs(n, seed=200 + case) * 2 - 0.5
        tex = dict(pixels=meshes.random_texture(16, 8, 3 + case % 2, seed=case), wrap=case % 2)
        a = refharness.run_internal(v, res, uvs=uv, texture=tex, strategy=strategy)
        b = oracle.voxelize(v, res, uvs=uv, texture=tex, strategy=strategy)
        assert np.array_equal(a["xyz"], b["xyz"])
        assert np.array_equal(a["wrgb"].view(np.uint32), b["wrgb"].view(np.uint32))


@pytest.mark.skipif(not refharness.available(), reason="oracle/_ref is only built where /root/reference exists")
@pytest.mark.parametrize("extent,resolution,supersampling", [(0.0005, 512, 1), (0.02, 128, 1), (0.004, 128, 2),
                                                              (0.002, 4096, 1), (0.002, 2688, 1)])
def test_oracle_matches_live_reference_on_the_inputs_of_the_gpu_classifier_tests(extent, resolution, supersampling):
    """The GPU tests of the two occupancy classifiers, of the job parts and of the > 65536-chunk grid compare against the
    oracle on micro-triangles, ordinary triangles with big axis-aligned leaves, 2x supersampling and a 2688^3 grid
    (tests/test_gpu_parity.py); here the oracle itself is pinned on the same inputs against the reference built from
    /root/reference (public API; the patched build where the down-scaling matters).  2688 = 42 chunks per axis is not a
    power of two: there the reference never dispatches the chunks whose Morton index is >= 42^3 (SURVEY B2,
    src/obj2voxel.cpp:503-505) while the oracle and the product cover the whole grid — the reference's voxels must be
    exactly the oracle's voxels inside the chunks it dispatched; 4096 pins the same input on a power-of-two grid."""
    big = resolution > 2048
    v = meshes.random_triangles(3000 if big else 20000, extent, seed=9 if big else 31)
    if not big:
        v = np.concatenate([v, meshes.unit_cube() * np.float32(0.9) + np.float32(0.05)])
    want = refharness.run_api(v, resolution, supersampling=supersampling, bounds=meshes.UNIT_BOUNDS,
                              patched_downscale=supersampling > 1)["voxels"]
    got = oracle.voxelize(v, resolution, supersampling=supersampling, bounds=meshes.UNIT_BOUNDS)["voxels"]
    chunks = (resolution * supersampling + 63) // 64
    if chunks & (chunks - 1) != 0:
        c = got[:, :3] // (64 // supersampling)
        morton = np.array([oracle.ileave3(int(x), int(y), int(z)) for x, y, z in c.tolist()], dtype=np.uint64)
        assert len(got) > len(want) > 0
        got = got[morton < np.uint64(chunks) ** 3]
    assert np.array_equal(got, want)
