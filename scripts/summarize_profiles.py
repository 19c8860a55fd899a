"""Turns gpurun_out/ ncu outputs into the committed summaries under profiles/ (run here, no GPU needed).
  python scripts/summarize_profiles.py launches <launches.csv> <out.md> "<title>"
  python scripts/summarize_profiles.py raw <report.ncu-rep> <out.md> "<title>"
  python scripts/summarize_profiles.py traffic <report.ncu-rep> <out.json> <workload>
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_write.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def launches(path, out, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0]
        name = "torch (input generation)" if ("at::" in name or "elementwise" in name) else name.replace("o2v::<unnamed>::", "").replace("void ", "")
        agg.setdefault(name, []).append(float(r[vi].replace(",", "")))
    ours = {k: v for k, v in agg.items() if not k.startswith("torch")}
    total = sum(sum(v) for v in ours.values())
    with open(out, "w") as f:
        f.write("# %s\n\nSource: `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare "
                "shares, not absolutes).\n\n| kernel | launches | ms / launch | share of our kernels |\n|---|---|---|---|\n" % title)
        for k, v in ours.items():
            f.write("| `%s` | %d | %.3f | %.1f %% |\n" % (k, len(v), sum(v) / len(v) / 1e6, 100 * sum(v) / total))
        f.write("\nTotal of our kernels per step: %.3f ms (over %d steps captured)\n" % (total / 1e6 / max(1, len(next(iter(ours.values())))), len(next(iter(ours.values())))))


def raw(rep, out, title):
    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(text.splitlines()))
    hdr, units = rows[0], dict(zip(rows[0], rows[1]))
    with open(out, "w") as f:
        f.write("# %s\n\nSource: `ncu --set full --clock-control none --import-source on`, read with `ncu -i … --page raw --csv`.\n\n" % title)
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % d.get("Kernel Name", "?").split("(")[0])
            for k in KEYS:
                if k in d and d[k] != "":
                    f.write("| %s | %s | %s |\n" % (k, d[k], units.get(k, "")))
            f.write("\n")


def traffic(rep, out, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of every kernel in the report -> JSON for bench.py."""
    import json

    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(text.splitlines()))
    hdr, units = rows[0], dict(zip(rows[0], rows[1]))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    kernels = {}
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        name = d.get("Kernel Name", "?").split("(")[0].split("::")[-1]
        total = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            total += float(d[k].replace(",", "")) * scale[units[k]]
        kernels[name] = {"dram_bytes_per_launch": int(total), "gpu_time_ms": float(d["gpu__time_duration.sum"]) *
                         {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units["gpu__time_duration.sum"]]}
    json.dump({"workload": workload, "source": rep.split("/")[-1] + " (ncu --set full, one launch per kernel)",
               "kernels": kernels}, open(out, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])
