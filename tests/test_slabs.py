"""CPU: Z-slab partitioning and the world_size-2 plumbing (gloo) of the multi-GPU path."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from obj2voxel_b200 import meshes, slabs
from oracle import oracle


@pytest.mark.parametrize("res,world", [(1024, 8), (2048, 8), (512, 2), (256, 4), (64, 2), (16, 2), (100, 3)])
def test_equal_slabs_cover_the_grid_once(res, world):
    b = slabs.equal_slabs(res, world)
    assert len(b) == world + 1 and b[0] == 0 and b[-1] >= res
    assert all(x % 8 == 0 for x in b) and all(b[i] <= b[i + 1] for i in range(world))
    if res >= 64 * world:
        assert all(x % 64 == 0 for x in b)  # whole reference chunk rows


def test_a_rank_without_rows_gets_no_slab():
    """More ranks than rows: the surplus rank must skip the job — (0, 0) would mean 'whole grid' to the engine."""
    b = slabs.equal_slabs(8, 2)
    assert b == [0, 0, 8]
    assert slabs.my_slab(b, 0) is None and slabs.my_slab(b, 1) == (0, 8)
    import obj2voxel_b200 as o2v
    with pytest.raises(ValueError):
        o2v.make_params(resolution=8, slab=(0, 0))


def test_balanced_slabs_follow_the_histogram():
    hist = np.zeros(16)
    hist[:4] = 100.0  # all the work is in the first quarter of the grid
    hist[4:] = 1.0
    b = slabs.balanced_slabs(hist, 1024, 4)
    assert b[0] == 0 and b[-1] == 1024 and all(x % 64 == 0 for x in b)
    assert b[1] <= 128 and b[3] <= 320
    assert slabs.balanced_slabs(np.zeros(16), 1024, 4) == slabs.equal_slabs(1024, 4)


def test_slab_union_equals_whole_grid_on_the_oracle():
    """Ownership by z range needs no voxel exchange: every voxel of the full result lies in exactly one slab and the
    per-slab results are the full result restricted to the slab (checked here on the CPU oracle; the GPU version of this
    property is in test_gpu_parity.py)."""
    v = meshes.random_triangles(400, 0.05, seed=31)
    full = oracle.voxelize(v, 128, strategy=1, bounds=[-0.1, -0.1, -0.1, 1.1, 1.1, 1.1])["voxels"]
    bounds = slabs.equal_slabs(128, 2)
    owner = np.searchsorted(bounds, full[:, 2], side="right") - 1
    assert set(owner.tolist()) <= {0, 1}
    assert (owner == 0).sum() + (owner == 1).sum() == len(full)


WORKER = r"""
import os, sys
sys.path.insert(0, os.environ["O2V_ROOT"])
import numpy as np, torch, torch.distributed as dist
from obj2voxel_b200 import meshes, slabs
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 1000
verts = torch.from_numpy(meshes.random_triangles(n, 0.01, seed=4)) if rank == 0 else torch.empty((n, 9), dtype=torch.float32)
slabs.broadcast_mesh([verts, None], src=0)
want = torch.from_numpy(meshes.random_triangles(n, 0.01, seed=4))
assert torch.equal(verts, want), "broadcast mismatch"
b = slabs.equal_slabs(512, world)
z0, z1 = slabs.my_slab(b, rank)
assert z1 > z0
total = slabs.allreduce_counts([z1 - z0, rank + 1], device="cpu")
assert total[0] == b[-1] and total[1] == world * (world + 1) // 2, total
dist.barrier()
dist.destroy_process_group()
open(os.path.join(os.environ["O2V_OUT"], "rank%d.ok" % rank), "w").write("ok")
"""


def test_world_size_two_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, O2V_ROOT=ROOT, O2V_OUT=str(tmp_path), OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert (tmp_path / "rank0.ok").exists() and (tmp_path / "rank1.ok").exists()  # stdout of the ranks interleaves
