#!/usr/bin/env python
"""bench.py — one "step" = one pass of the voxelization hot path over one batch of synthetic triangles.

Workload (N = 1 and N > 1): BASELINE.json config 4 — the configuration its metric ("Mtri/s & Mvoxel/s @1024^3") is quoted
on: 10 M random triangles (centres U[0,1]^3, vertex offsets U[+-0.001]^3, splitmix64 stream), resolution 1024 with 2x
supersampling (sample grid 2048^3), MAX strategy, mesh bounds given.  With N GPUs the sample grid is split into N Z-slabs
of whole 64-voxel chunk rows (strong scaling: the total work is fixed); every rank holds the whole triangle array (one
broadcast outside the timed region), owns its slab's voxels, and no voxel data is exchanged — the only collective in a
step is the barrier/all-reduce used for timing and the per-slab counts.

JSON line (rank 0): value = whole-job triangles/s with inputs resident in HBM (CUDA events, max over ranks);
e2e = the same through the reference-facing C API with HOST buffers (H2D + kernels + D2H inside the timed region);
roofline = algorithmic bytes of the dominant kernel / its CUDA-event duration vs the measured HBM peak;
cpu_baseline = the unmodified reference's threaded CPU path (oracle/_ref) on a bounded sample of the same workload.
`--impl reference` times that CPU path alone and prints the same line with "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (config key in meshes.CONFIGS, human description)
    "cfg4": "BASELINE config 4: 10M random triangles (offsets +-0.001), res 1024, 2x supersampling (sample 2048^3), MAX",
    "r1024": "10M random triangles (offsets +-0.001), res 1024, no supersampling, MAX",
    "cfg3": "BASELINE config 3: 1M random UV-textured triangles (offsets +-0.004), res 512, BLEND",
    "cfg5": "BASELINE config 5: 100M sub-voxel micro-triangles, res 2048, MAX",
    "cfg2": "BASELINE config 2: ~70k-triangle lumpy sphere, res 256, MAX",
}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profiled_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu capture."""
    path = os.path.join(ROOT, "profiles", "voxelize_traffic.json")
    if os.path.exists(path):
        try:
            data = json.load(open(path))
            if data.get("workload") == workload:
                return data.get("kernels", {}).get(kernel, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        """Start of the timed region: only later samples count."""
        self.t0 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0 = getattr(self, "t0", 0.0)
        t1 = time.time()
        for stamp, line in self.lines:
            if stamp < t0 or stamp > t1:
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[2:6]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def workload_spec(name):
    from obj2voxel_b200 import meshes

    cfg = dict(meshes.CONFIGS[name])
    if cfg["kind"] == "random":
        e = cfg["extent"]
        cfg["bounds"] = [-e, -e, -e, 1.0 + e, 1.0 + e, 1.0 + e]
    else:
        cfg["bounds"] = None
    return cfg


def config_dict(workload, cfg, n_tri, world):
    """`config` of the JSON line — the same dict from both arms (ours and --impl reference) for a given workload and N."""
    tri_bytes = 64 if cfg.get("textured") else 36
    return {"workload": workload, "description": WORKLOADS[workload], "triangles": int(n_tri),
            "resolution": cfg["resolution"], "supersampling": cfg["supersampling"],
            "strategy": "blend" if cfg["strategy"] else "max",
            "partition": "whole grid on one GPU" if world == 1 else
                         "%d Z-slabs of whole 64-voxel chunk rows, one per GPU (strong scaling)" % world,
            "l2": ("inputs (%d MB of triangles) exceed the 126 MB L2; no explicit flush" if n_tri * tri_bytes > 126e6 else
                   "inputs (%d MB of triangles) fit the 126 MB L2 and no flush is done: a side measurement, not the "
                   "headline") % (n_tri * tri_bytes // 1000000)}


def golden_checksums():
    """Voxel count + order-independent record hash of the UNMODIFIED reference's output per workload
    (tests/golden/full_size_checksums.json, written by tests/golden/make_full_size_checksums.py from oracle/_ref)."""
    path = os.path.join(ROOT, "tests", "golden", "full_size_checksums.json")
    try:
        return json.load(open(path))
    except Exception:
        return {}


def host_mesh(cfg, count=None):
    from obj2voxel_b200 import meshes

    if cfg["kind"] == "sphere":
        return meshes.lumpy_sphere(), None
    n = cfg["n"] if count is None else count
    verts = meshes.random_triangles(n, cfg["extent"], seed=1)
    uvs = meshes.random_uvs(n, seed=2) if cfg.get("textured") else None
    return verts, uvs


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference (oracle/_ref) on the host cores

def reference_sample_run(cfg, sample, workers):
    """Times the reference's own threaded CPU path on the first `sample` triangles of the workload.  Supersampling is
    timed as resolution*ss with ss=1: identical work up to the reference's broken downscale."""
    from obj2voxel_b200 import meshes
    from oracle import refharness

    verts, uvs = host_mesh(cfg, sample)
    texture = dict(pixels=meshes.random_texture(256, 256, 3), wrap=1) if uvs is not None else None
    res = cfg["resolution"] * cfg["supersampling"]
    r = refharness.run_api(verts, res, uvs=uvs, texture=texture, strategy=cfg["strategy"], bounds=cfg["bounds"],
                           workers=workers, collect=False)
    return r["seconds"], r["count"], len(verts)


def cpu_baseline(cfg, target_seconds=12.0):
    from oracle import refharness

    if not refharness.available():
        return port_baseline(cfg)
    workers = refharness.hardware_threads()
    total = cfg.get("n", 70000)
    sample = min(total, 100_000)
    secs, voxels, n = reference_sample_run(cfg, sample, workers)
    if secs < target_seconds / 3 and sample < total:
        sample = int(min(total, max(sample, sample * target_seconds / max(secs, 1e-3))))
        secs, voxels, n = reference_sample_run(cfg, sample, workers)
    return {"value": n / secs / 1e6, "unit": "Mtri/s", "cores": workers, "kind": "reference",
            "mvoxel_per_s": voxels / secs / 1e6, "seconds": secs,
            "sample": "first %d of %d triangles of the same workload, obj2voxel_voxelize() wall clock with %d worker "
                      "threads (reference built from /root/reference, -O3, no FMA)" % (n, total, workers)}


def port_baseline(cfg):
    from oracle import oracle

    sample = min(cfg.get("n", 70000), 50_000)
    verts, _ = host_mesh(cfg, sample)
    t0 = time.time()
    r = oracle.voxelize(verts, cfg["resolution"] * cfg["supersampling"], strategy=cfg["strategy"], bounds=cfg["bounds"],
                        downscale=False)
    secs = time.time() - t0
    return {"value": len(verts) / secs / 1e6, "unit": "Mtri/s", "cores": os.cpu_count(), "kind": "port",
            "mvoxel_per_s": len(r["xyz"]) / secs / 1e6, "seconds": secs,
            "sample": "first %d triangles, oracle/liboracle.so (C restatement), all cores" % len(verts)}


def run_reference_arm(args, cfg, workload):
    from oracle import refharness

    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    workers = refharness.hardware_threads() if refharness.available() else os.cpu_count()
    total = cfg.get("n", 70000)
    # calibrate the per-step sample so the whole --steps/--warmup run ends within a few minutes (~4 s per step)
    sample = min(total, 50_000)
    if refharness.available():
        secs, _, _ = reference_sample_run(cfg, sample, workers)
        sample = int(min(total, max(10_000, sample * 4.0 / max(secs, 1e-3))))
        times, voxels = [], 0
        for step in range(args.warmup + args.steps):
            secs, voxels, n = reference_sample_run(cfg, sample, workers)
            if step >= args.warmup:
                times.append(secs)
        kind = "reference"
    else:
        base = port_baseline(cfg)
        times, voxels, n, kind = [base["seconds"]], 0, sample, "port"
    mean = float(np.mean(times))
    value = sample / mean / 1e6
    line = {
        "impl": "reference", "metric": "triangles_per_second", "value": value, "unit": "Mtri/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(workload, cfg, total, args.gpus),
        "sample": "each step: the first %d of the %d triangles at resolution %d x supersampling 1 (the reference's "
                  "downscale is broken, BASELINE.md section 3: same work per triangle without it)" %
                  (sample, total, cfg["resolution"] * cfg["supersampling"]),
        "mvoxel_per_s": voxels / mean / 1e6,
        "cpu_baseline": {"value": value, "unit": "Mtri/s", "cores": workers, "kind": kind,
                         "sample": "each step: first %d of %d triangles of the workload through obj2voxel_voxelize() "
                                   "with %d worker threads" % (sample, total, workers)},
        "e2e": {"value": value, "unit": "Mtri/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# our arm

class Leg:
    """One workload on this rank's GPU: mesh resident in HBM (generated in place; with N > 1 ranks on the occupancy-only
    path the rank keeps only the triangles that can reach its Z-slab — distributed once at ingest, outside any timed
    region), parameters, and the reference's checksum to compare with."""

    def __init__(self, engine, name, cfg, rank, world, device, broadcast=False):
        import torch
        import torch.distributed as dist

        import obj2voxel_b200 as o2v
        from obj2voxel_b200 import meshes, slabs

        self.engine, self.name, self.cfg, self.rank, self.world, self.device = engine, name, cfg, rank, world, device
        if cfg["kind"] == "sphere":
            verts = torch.from_numpy(meshes.lumpy_sphere()).to(device)
            uvs = None
        elif broadcast and world > 1:
            # the ingest rank owns the mesh and broadcasts it once (NCCL over NVLink)
            n = cfg["n"]
            if rank == 0:
                verts = meshes.random_triangles_torch(n, cfg["extent"], seed=1, device=device)
            else:
                verts = torch.empty((n, 9), dtype=torch.float32, device=device)
            uvs = None
            if cfg.get("textured"):
                uvs = meshes.random_uvs_torch(n, seed=2, device=device) if rank == 0 else \
                    torch.empty((n, 6), dtype=torch.float32, device=device)
            slabs.broadcast_mesh([verts, uvs], src=0)
        else:
            verts = meshes.random_triangles_torch(cfg["n"], cfg["extent"], seed=1, device=device)
            uvs = meshes.random_uvs_torch(cfg["n"], seed=2, device=device) if cfg.get("textured") else None
        self.n_tri = int(verts.shape[0])
        self.textures = []
        if uvs is not None:
            self.textures = [(torch.from_numpy(meshes.random_texture(256, 256, 3)).to(device), o2v.UV_WRAP)]
        S = cfg["resolution"] * cfg["supersampling"]
        self.bounds = slabs.equal_slabs(S, world)
        slab = slabs.my_slab(self.bounds, rank)
        self.empty = world > 1 and slab is None  # more ranks than rows: this rank owns nothing and skips the steps
        z0, z1 = slab if slab is not None else (0, 0)
        kw = dict(resolution=cfg["resolution"], supersampling=cfg["supersampling"], strategy=cfg["strategy"],
                  bounds=cfg["bounds"])
        self.kw = kw
        self.full_verts, self.full_uvs = verts, uvs
        self.verts, self.uvs = verts, uvs
        self.params = o2v.make_params(slab=(z0, z1) if world > 1 and not self.empty else None, **kw)
        self.distributed_once = False
        if world > 1 and uvs is None and not self.empty and cfg["bounds"] is not None:
            # z-binned distribution: this rank keeps what can reach its slab (o2v_b200_filter_slab)
            self.verts = engine.filter_slab(verts, self.params)
            self.params = o2v.make_params(slab=(z0, z1), slab_filtered=1, **kw)
            self.distributed_once = True
        self.tri_bytes = 64 if uvs is not None else 36

    def step(self, params=None):
        if self.empty:
            return None
        return self.engine.voxelize_device(self.verts, params or self.params, uvs=self.uvs, textures=self.textures)

    def result_check(self):
        """(voxel count, record hash) of the last step on this rank, hash computed on the device."""
        if self.empty:
            return 0, 0
        return self.engine.result_count(), self.engine.result_hash()


def gather_check(leg, golden, dist_on):
    """Sums the per-rank voxel counts and record hashes (mod 2^64) and compares with the reference's."""
    import torch
    import torch.distributed as dist

    count, h = leg.result_check()
    parts = [(count, h)]
    if dist_on:
        t = torch.tensor([count, h - (1 << 64) if h >= (1 << 63) else h], dtype=torch.int64, device=leg.device)
        out = [torch.empty_like(t) for _ in range(leg.world)]
        dist.all_gather(out, t)
        parts = [(int(o[0].item()), int(o[1].item()) & ((1 << 64) - 1)) for o in out]
    total = sum(p[0] for p in parts)
    digest = sum(p[1] for p in parts) & ((1 << 64) - 1)
    ref = golden.get(leg.name)
    ok = None
    if ref is not None and "hash64" in ref:
        ok = bool(total == ref["voxels"] and digest == ref["hash64"])
    return total, digest, ok


def time_leg(leg, steps, warmup, sync_all, params=None):
    """Device-resident steps: CUDA events on the stream the kernels run on; returns (elapsed ms over `steps`, last stats,
    per-step kernel timings)."""
    import torch

    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stats = None
    for _ in range(warmup):
        stats = leg.step(params)
    sync_all()
    rows = []
    start.record()
    for _ in range(steps):
        stats = leg.step(params)
        if stats is not None:
            rows.append(stats)
    stop.record()
    sync_all()
    return start.elapsed_time(stop), stats, rows


def side_config(engine, name, rank, world, device, sync_all, golden, steps):
    """One of BASELINE.json's other named configs, measured like the headline (device-resident, max over ranks) and
    checked against the reference's checksum; reported under `configs`, never as the headline."""
    import torch
    import torch.distributed as dist

    cfg = workload_spec(name)
    leg = Leg(engine, name, cfg, rank, world, device)
    elapsed, stats, _rows = time_leg(leg, steps, 3, sync_all)
    total, digest, ok = gather_check(leg, golden, world > 1)
    t = torch.tensor([elapsed], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    out = {"mtri_per_s": leg.n_tri / (ms * 1e-3) / 1e6, "mvoxel_per_s": total / (ms * 1e-3) / 1e6, "ms_per_step": ms,
           "triangles": leg.n_tri, "voxels": total, "hash64": digest, "hash_ok": ok, "steps": steps,
           "path": "occupancy-only" if (stats is not None and stats["occupancy_path"]) else "weighted fold",
           "config": config_dict(name, cfg, leg.n_tri, world)}
    del leg
    torch.cuda.empty_cache()
    return out


def run_ours(args, cfg, workload):
    import torch
    import torch.distributed as dist

    import obj2voxel_b200 as o2v
    from obj2voxel_b200 import meshes, slabs

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    distributed = world > 1
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if distributed:
        dist.init_process_group("nccl", device_id=device)

    engine = o2v.Engine(local)  # raises without a GPU: there is no fallback
    golden = golden_checksums()

    def sync_all():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- inputs resident in HBM; the ingest rank broadcasts the triangle array once, every rank keeps its slab's
    # triangles (not timed) ----
    leg = Leg(engine, workload, cfg, rank, world, device, broadcast=True)
    n_tri = leg.n_tri
    uvs = leg.uvs

    # ---- device-resident steps ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # started before the warm-up (nvidia-smi takes a while to come up); samples are windowed
    time_leg(leg, 0, args.warmup, sync_all)
    sampler.mark()
    elapsed_ms, stats, rows = time_leg(leg, args.steps, 0, sync_all)
    clocks = sampler.stop() if rank == 0 else None
    if stats is None:  # a rank without rows
        stats = {k: 0 for k in ("voxels", "contributions", "clip_calls", "leaves", "light_tiles", "heavy_tiles",
                                "occupancy_path", "slab_triangles")}
    kernel_ms = [r["ms_voxelize"] for r in rows] or [0.0]
    setup_ms = [r["ms_setup"] for r in rows] or [0.0]
    clip_ms = [r["ms_clip"] for r in rows] or [0.0]
    classify_ms = [r["ms_classify"] for r in rows] or [0.0]
    launches = sum(r["kernel_launches"] for r in rows)
    occupancy_path = bool(stats["occupancy_path"])
    voxels, digest, hash_ok = gather_check(leg, golden, distributed)
    if hash_ok is False:
        raise RuntimeError("%s: %d voxels, record hash %d differ from the reference's (%s)" %
                           (workload, voxels, digest, golden.get(workload)))

    # The same workload with the occupancy-only path switched off (weights and colours folded for every voxel): what a
    # coloured / textured mesh of this shape costs.  Reported beside the headline, not as the headline.
    weighted = None
    if occupancy_path and not distributed:
        wparams = o2v.make_params(occupancy_path=0, **leg.kw)
        wsteps = max(1, min(args.steps, 5))
        wms, wstats, _ = time_leg(leg, wsteps, 2, sync_all, params=wparams)
        wms /= wsteps
        wcount, wdigest = leg.result_check()
        if wcount != voxels or wdigest != digest:
            raise RuntimeError("weighted path: %d voxels / hash %d, occupancy path %d / %d" %
                               (wcount, wdigest, voxels, digest))
        weighted = {"value": n_tri / (wms * 1e-3) / 1e6, "unit": "Mtri/s", "ms_per_step": wms,
                    "clip_calls": wstats["clip_calls"], "contributions": wstats["contributions"],
                    "ms_clip_kernel": wstats["ms_clip"], "hash_ok": hash_ok,
                    "note": "same workload with occupancy_path=0: every (triangle, voxel) weight folded in reference "
                            "order (what coloured / textured meshes cost); identical records (same hash)"}

    counts = [stats["contributions"], stats["clip_calls"], stats["leaves"], launches]
    tile_split = (stats["light_tiles"], stats["heavy_tiles"])
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        counts = slabs.allreduce_counts(counts, device)
    elapsed_ms = float(t.item())
    contributions, clip_calls, leaves, launches = counts
    ms_per_step = elapsed_ms / args.steps
    rank_tri = int(stats["slab_triangles"]) if occupancy_path else n_tri
    rank_voxels = int(stats["voxels"])

    # ---- BASELINE.json's other named configs, same measurement, each checked against the reference's checksum ----
    pinned_verts = pinned_uvs = None
    if rank == 0:
        pinned_verts = torch.empty(leg.full_verts.shape, dtype=torch.float32, pin_memory=True)
        pinned_verts.copy_(leg.full_verts)
        if leg.full_uvs is not None:
            pinned_uvs = torch.empty(leg.full_uvs.shape, dtype=torch.float32, pin_memory=True)
            pinned_uvs.copy_(leg.full_uvs)
    leg_bounds, tri_bytes_e2e = leg.bounds, leg.tri_bytes
    del leg
    torch.cuda.empty_cache()
    side = {}
    side_steps = max(1, min(args.steps, 5))
    for name in ("cfg2", "cfg3", "cfg5"):
        if name != workload and os.environ.get("O2V_BENCH_SIDE", "1") != "0":
            side[name] = side_config(engine, name, rank, world, device, sync_all, golden, side_steps)

    # ---- end to end with HOST buffers: one process, all N devices (rank 0); the other ranks release theirs first ----
    engine.close()
    torch.cuda.empty_cache()
    sync_all()
    cpu_group = dist.new_group(backend="gloo") if distributed else None
    e2e = None
    if rank == 0:
        e2e = run_e2e(args, cfg, pinned_verts.numpy(), None if pinned_uvs is None else pinned_uvs.numpy(), n_tri,
                      tri_bytes_e2e, rank, world, voxels, golden, workload)
    if distributed:
        dist.barrier(group=cpu_group)  # on the CPU: no kernel of the waiting ranks sits on the devices rank 0 is driving

    if rank == 0:
        peak, peak_source = measured_peak()
        tri_bytes = 64 if uvs is not None else 36  # algorithmic read per triangle
        # dominant kernel of the step, by its live CUDA-event duration: the SAT classification kernel on the occupancy-only
        # path, the exact clip otherwise; one launch processes this rank's whole slab
        k_ms = float(np.mean(kernel_ms))
        c_ms = float(np.mean(clip_ms))
        f_ms = float(np.mean(classify_ms))
        dominant, d_ms = ("occupancyClassifyKernel", f_ms) if f_ms > c_ms else \
            ("occupancyClipKernel" if occupancy_path else "sparseClipKernel", c_ms)
        # algorithmic bytes of ONE launch on THIS rank: its slab's triangles (what the z-binned ingest left it) and the
        # voxels it emits
        alg_bytes = 16 * rank_voxels + tri_bytes * rank_tri
        achieved = alg_bytes / (d_ms * 1e-3) / 1e9 if d_ms > 0 else 0.0
        traffic = profiled_traffic(workload, dominant) if world == 1 else None
        baseline = cpu_baseline(cfg) if world == 1 else None
        line = {
            "metric": "triangles_per_second", "value": n_tri / (ms_per_step * 1e-3) / 1e6, "unit": "Mtri/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(workload, cfg, n_tri, world),
            "slab_bounds": leg_bounds if distributed else None,
            "mvoxel_per_s": voxels / (ms_per_step * 1e-3) / 1e6, "voxels": voxels, "hash64": digest, "hash_ok": hash_ok,
            "contributions": contributions,
            "clip_calls": clip_calls, "leaves": leaves, "light_tiles_rank0": tile_split[0],
            "heavy_tiles_rank0": tile_split[1],
            "ms_setup_rank0": float(np.mean(setup_ms)), "ms_voxelize_rank0": k_ms, "ms_clip_kernel_rank0": c_ms,
            "ms_classify_kernel_rank0": f_ms,
            "path": "occupancy-only (every triangle MATERIALLESS: output colour is white whatever the weights)"
                    if occupancy_path else "weighted fold",
            "hbm_write_gbs": 16 * voxels / (ms_per_step * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_source, "rank0_triangles": rank_tri, "rank0_voxels": rank_voxels,
                         "note": "rank 0: algorithmic bytes = 16 B x its voxels + %d B x its slab's triangles per launch / "
                                 "CUDA-event duration of its dominant kernel; traffic is an ncu capture at N = 1 only"
                                 % tri_bytes},
            "e2e": e2e,
            "configs": side,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if weighted is not None:
            line["weighted_path"] = weighted
        if baseline is not None:
            line["cpu_baseline"] = baseline
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier(group=cpu_group)
        dist.destroy_process_group()
    return 0


def run_e2e(args, cfg, host_verts, host_uvs, n_tri, tri_bytes, rank, world, voxels, golden, workload):
    """End to end with HOST buffers through the reference-facing C API: obj2voxel instance, bulk triangle input (pinned host
    memory), obj2voxel_voxelize(), voxel callback receiving reference-layout quads.  H2D, kernels, D2H and the host-side
    expansion of downloaded bitmaps are all inside the timed region.  With N GPUs this is ONE process driving N devices
    (obj2voxel_b200_set_devices: Z-slabs, triangles exchanged over peer memory) — rank 0 runs it while the other ranks
    have released their devices.  Also timed at N = 1: the same job from a pageable array and through the reference's
    one-callback-per-triangle ingestion."""
    import ctypes as C

    import obj2voxel_b200 as o2v
    from obj2voxel_b200 import _lib, meshes

    lib = o2v.load()
    lib.obj2voxel_set_log_level(_lib.LOG_ERROR)
    tex_obj = o2v.Texture(meshes.random_texture(256, 256, 3), wrap=o2v.UV_WRAP) if host_uvs is not None else None
    devices = list(range(world))
    e2e_steps = max(1, min(args.steps, 5))

    def one_job(verts, mode="bulk", collect=False):
        inst = o2v.Instance()
        source = None
        if mode == "callback":
            source = _lib.ArraySource(verts.ctypes.data, len(verts), 0)
            lib.obj2voxel_set_input_callback(inst.handle, C.cast(lib.obj2voxel_b200_array_source_next,
                                                                 _lib.TRIANGLE_CALLBACK), C.addressof(source))
        else:
            inst.set_input_triangles(verts, uvs=host_uvs, texture=tex_obj)
        received = {"n": 0, "chunks": []}
        counter = _lib.CountingSink(0, 0)
        if collect:
            # untimed verification run: a Python callback that keeps a copy of every batch
            def on_voxels(_data, quads, count, received=received):
                received["n"] += count
                if count:
                    received["chunks"].append(np.ctypeslib.as_array(quads, shape=(count, 4)).copy())
                return True

            cb = _lib.VOXEL_CALLBACK(on_voxels)
        else:
            # timed runs: the C callback a C caller would pass (it counts); a Python callback would add the
            # interpreter's cost per call to a path whose calls come from the library's own threads
            cb = C.cast(lib.obj2voxel_b200_counting_sink_write, _lib.VOXEL_CALLBACK)
        lib.obj2voxel_set_output_callback(inst.handle, cb, C.addressof(counter))
        inst.set_resolution(cfg["resolution"])
        inst.set_supersampling(cfg["supersampling"])
        inst.set_color_strategy(cfg["strategy"])
        if cfg["bounds"] is not None:
            inst.set_mesh_boundaries(cfg["bounds"])
        inst.set_devices(devices)
        t0 = time.perf_counter()
        err = inst.voxelize()
        dt = time.perf_counter() - t0
        stats = inst.stats()
        inst.free()
        if err != 0:
            raise RuntimeError("obj2voxel_voxelize failed with error %d" % err)
        if not collect:
            received["n"] = int(counter.voxels)
        return dt, received, stats

    def timed(verts, mode="bulk", steps=e2e_steps):
        times, got = [], 0
        for step in range(1 + steps):
            dt, received, _ = one_job(verts, mode)
            if step > 0:
                times.append(dt)
                got = received["n"]
        if got != voxels:
            raise RuntimeError("e2e (%s) voxel count %d != device-resident count %d" % (mode, got, voxels))
        return float(np.mean(times))

    seconds = timed(host_verts)
    # the records the callback received, once, untimed: same order-independent hash as the reference's output
    _, received, stats = one_job(host_verts, collect=True)
    records = np.concatenate(received["chunks"]) if received["chunks"] else np.zeros((0, 4), np.uint32)
    digest = meshes.record_hash(records)
    ref = golden.get(workload)
    hash_ok = bool(len(records) == ref["voxels"] and digest == ref["hash64"]) if ref and "hash64" in ref else None
    del records, received
    if hash_ok is False:
        raise RuntimeError("e2e records differ from the reference's (hash %d)" % digest)
    packed = bool(stats["download_bytes"] < 16 * voxels)
    out = {"value": n_tri / seconds / 1e6, "unit": "Mtri/s", "ms_per_step": seconds * 1e3,
           "h2d_bytes_per_step": int(n_tri * tri_bytes),
           "d2h_bytes_per_step": int(stats["download_bytes"]), "voxel_callback_bytes_per_step": int(16 * voxels),
           "hash_ok": hash_ok, "devices": world,
           "api": "obj2voxel_b200_set_input_triangles (pinned host array) + obj2voxel_b200_set_devices(%d) + "
                  "obj2voxel_voxelize + voxel callback; %s" %
                  (world, "the result crosses PCIe as packed positions (%d bytes per voxel) and host threads write the "
                          "quads the callback receives%s" % (round(stats["download_bytes"] / max(voxels, 1)),
                                                            "; on one device the triangle array goes up in pieces and "
                                                            "every piece is voxelized while the next one uploads"
                                                            if world == 1 else "") if packed else
                   "records cross PCIe as they are, each device's over its own link")}
    if world == 1 and host_uvs is None:
        pageable = np.array(host_verts, copy=True)  # plain numpy memory
        ms = timed(pageable, steps=min(e2e_steps, 3)) * 1e3
        out["pageable_input"] = {"ms_per_step": ms, "value": n_tri / (ms * 1e-3) / 1e6, "unit": "Mtri/s",
                                 "note": "same job from a pageable array: host threads stage it through pinned buffers"}
        ms = timed(host_verts, mode="callback", steps=1) * 1e3
        out["callback_input"] = {"ms_per_step": ms, "value": n_tri / (ms * 1e-3) / 1e6, "unit": "Mtri/s",
                                 "note": "same job through obj2voxel_set_input_callback: one indirect call per triangle "
                                         "(src/obj2voxel.cpp:585-588), then the same device path"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg = workload_spec(args.workload)
    if args.impl == "reference":
        return run_reference_arm(args, cfg, args.workload)
    return run_ours(args, cfg, args.workload)


if __name__ == "__main__":
    sys.exit(main())
