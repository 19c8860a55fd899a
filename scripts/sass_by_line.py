"""Joins the per-instruction counters of an `ncu --set full --import-source on` report (SASS page) with the line table of
the same build (`nvdisasm -g`) and prints, per source line of a kernel, the share of executed warp instructions, the
average number of active lanes and the share of stall samples.  Runs here (no GPU):
  python scripts/sass_by_line.py <report.ncu-rep> <kernel name substring> [<cubin name substring, default o2v_occupancy> [<mangled name substring>]]
The report and the library must come from the same build (same instruction count, checked)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def profiled(rep, kernel):
    text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel],
                          capture_output=True, text=True).stdout
    rows = list(csv.reader(text.splitlines()))
    h = rows[1]
    ia, isamp, iex, ith = (h.index(k) for k in ("Source", "# Samples", "Instructions Executed", "Avg. Threads Executed"))
    data = [r for r in rows[2:] if len(r) > 10 and r[isamp].isdigit()]
    return [(int(r[isamp]), int(r[iex]), float(r[ith]), r[ia]) for r in data]


def line_table(kernel, cubin_hint):
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "obj2voxel_b200", "libobj2voxel_b200.so")],
                       cwd=tmp, capture_output=True)
        cubin = [f for f in os.listdir(tmp) if cubin_hint in f][0]
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True,
                             text=True).stdout.splitlines()
    start = [i for i, l in enumerate(txt) if l.startswith(".text.") and kernel in l][0]
    end = [i for i, l in enumerate(txt) if i > start and l.strip().startswith(".section")][0]
    out, cur = [], ("?", 0)
    for l in txt[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        elif re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
            out.append(cur)
    return out


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    hint = sys.argv[3] if len(sys.argv) > 3 else "o2v_occupancy"
    mangled = sys.argv[4] if len(sys.argv) > 4 else kernel  # e.g. sparseFoldKernelILb1 for the <true> instance
    prof, lines = profiled(rep, kernel), line_table(mangled, hint)
    if len(prof) == 2 * len(lines):  # some captures list the kernel twice
        prof = prof[:len(lines)]
    if len(prof) != len(lines):
        sys.exit("report (%d instructions) and library (%d) are not the same build" % (len(prof), len(lines)))
    total = sum(p[1] for p in prof)
    samples = sum(p[0] for p in prof)
    ex, lanes, st = collections.Counter(), collections.Counter(), collections.Counter()
    for key, (s, e, t, _) in zip(lines, prof):
        ex[key] += e
        lanes[key] += e * t
        st[key] += s
    source = {}
    print("# `%s`: executed warp instructions by source line\n" % kernel)
    print("Source: %s (SASS page) joined with `nvdisasm -g` of the same build; %.3f G warp instructions, %d stall "
          "samples.\n" % (os.path.basename(rep), total / 1e9, samples))
    print("| file:line | % of executed instructions | active lanes | % of stall samples | source |\n|---|---|---|---|---|")
    for (f, ln), e in ex.most_common(int(os.environ.get("O2V_LINES", "30"))):
        if f not in source:
            path = os.path.join(ROOT, "obj2voxel_b200", "csrc", f)
            source[f] = open(path).read().splitlines() if os.path.exists(path) else []
        text = source[f][ln - 1].strip()[:80].replace("|", "\\|") if 0 < ln <= len(source[f]) else ""
        print("| %s:%d | %.1f | %.1f | %.1f | `%s` |" % (f, ln, 100 * e / total, lanes[(f, ln)] / max(e, 1),
                                                       100 * st[(f, ln)] / max(samples, 1), text))


if __name__ == "__main__":
    main()
