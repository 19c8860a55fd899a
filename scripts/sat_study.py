"""CPU study of two leads for the classify kernel, on the fuzz regimes of tests/test_sat_classifier.py:
  * `scaled`: every edge function divided by its `certain` threshold, the nine values of a voxel reduced to one minimum
    (12 compares per voxel become 5 min + 2 compares);
  * a smaller `certain` shrink (1/64 instead of 1/32) where the sample resolution allows it.
Prints, per regime and variant, the verdict counts and the two kinds of violation (both must be 0 for a variant to be
usable).  Nothing here is a product path: the verdict variants live in the test shim (o2v_hostmath_test.cpp)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_sat_classifier as regimes  # noqa: E402

fp = C.POINTER(C.c_float)
lib = C.CDLL(os.path.join(ROOT, "obj2voxel_b200", "libo2v_hostmath_test.so"))
lib.o2vt_classify_study.argtypes = [fp, C.c_size_t, C.c_ulonglong, C.c_float, C.c_int, C.POINTER(C.c_ulonglong)]


def run(leaves, margin, scaled, max_volume=200_000):
    leaves = np.ascontiguousarray(leaves, np.float32).reshape(-1, 9)
    out = (C.c_ulonglong * 8)()
    lib.o2vt_classify_study(leaves.ctypes.data_as(fp), len(leaves), max_volume, margin, scaled, out)
    keys = ["pairs", "miss", "uncertain", "certain", "hits", "miss_but_hit", "certain_but_no_hit", "skipped"]
    return dict(zip(keys, [int(x) for x in out]))


def main():
    variants = [("header form, 1/32", 1 / 32, 0), ("scaled, 1/32", 1 / 32, 1), ("header form, 1/64", 1 / 64, 0),
                ("scaled, 1/64", 1 / 64, 1), ("scaled, 1/128", 1 / 128, 1)]
    print("%-22s %-20s %10s %10s %10s %10s %6s %6s" % ("regime", "variant", "pairs", "uncertain", "certain", "hits",
                                                     "m&hit", "c&!hit"))
    for name, make in regimes.REGIMES:
        leaves = make(np.random.default_rng(sum(map(ord, name))))
        for label, margin, scaled in variants:
            s = run(leaves, margin, scaled)
            print("%-22s %-20s %10d %10d %10d %10d %6d %6d" % (name, label, s["pairs"], s["uncertain"], s["certain"],
                                                               s["hits"], s["miss_but_hit"], s["certain_but_no_hit"]),
                  flush=True)


if __name__ == "__main__":
    main()
