import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def oracle_kwargs(g, downscale=True):
    """Arguments for oracle.oracle.voxelize from a golden fixture."""
    kw = dict(strategy=int(g["strategy"]), supersampling=int(g["supersampling"]), downscale=downscale)
    if "bounds" in g:
        kw["bounds"] = g["bounds"].tolist()
    if "unit" in g:
        kw["unit"] = g["unit"].tolist()
    if "uvs" in g:
        kw["uvs"] = g["uvs"]
        kw["texture"] = dict(pixels=g["tex_pixels"], wrap=int(g["tex_wrap"]))
    if "types" in g:
        kw["types"] = g["types"]
        kw["colors"] = g["colors"]
    return kw


def gpu_run(engine, g, **overrides):
    """Runs a golden fixture's input through the C-ABI host path; returns (sorted voxels, stats)."""
    import obj2voxel_b200 as o2v

    kw = dict(resolution=int(g["resolution"]), strategy=int(g["strategy"]), supersampling=int(g["supersampling"]))
    if "bounds" in g:
        kw["bounds"] = g["bounds"].tolist()
    if "unit" in g:
        kw["unit"] = g["unit"].tolist()
    kw.update(overrides)
    params = o2v.make_params(**kw)
    textures = [(g["tex_pixels"], int(g["tex_wrap"]))] if "uvs" in g else []
    voxels, stats = engine.voxelize_host(g["verts"], params, uvs=g.get("uvs"), types=g.get("types"),
                                         colors=g.get("colors"), textures=textures)
    return o2v.sort_voxels(voxels), stats


@pytest.fixture(scope="session")
def engine():
    import obj2voxel_b200 as o2v

    e = o2v.Engine(0)
    yield e
    e.close()
