"""Golden fixture for real-mesh parity of the OBJ reader: runs the UNMODIFIED reference (oracle/_ref, its own
tinyobjloader-based reader, src/io.cpp:194-312) on tinyobjloader/models/cornell_box.obj at resolution 64 and stores the
model's text (OBJ + MTL: test data, ~2.5 KB), the voxels and the per-colour counts in tests/golden/models/cornell_box_r64.npz.
CPU only; run in the build container: `python tests/golden/make_golden_models.py`."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refharness as R  # noqa: E402

MODELS = "/root/reference/tinyobjloader/models"


def main():
    cwd = os.getcwd()
    os.chdir(MODELS)  # tinyobjloader resolves mtllib relative to the working directory
    try:
        voxels = R.run_file("cornell_box.obj", 64)
    finally:
        os.chdir(cwd)
    colours, counts = np.unique(voxels[:, 3], return_counts=True)
    print(len(voxels), dict(zip([hex(int(c)) for c in colours], counts.tolist())))
    assert len(voxels) == 25574  # 17 639 white / 3 970 red / 3 965 green
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "models", "cornell_box_r64.npz"),
                        obj=np.frombuffer(open(os.path.join(MODELS, "cornell_box.obj"), "rb").read(), dtype=np.uint8),
                        mtl=np.frombuffer(open(os.path.join(MODELS, "cornell_box.mtl"), "rb").read(), dtype=np.uint8),
                        voxels=voxels, resolution=np.int64(64))


if __name__ == "__main__":
    main()
