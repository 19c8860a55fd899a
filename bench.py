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
        "config": {"workload": workload, "description": WORKLOADS[workload]},
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
    S = cfg["resolution"] * cfg["supersampling"]

    # ---- inputs resident in HBM; the ingest rank broadcasts the triangle array once (not timed) ----
    if cfg["kind"] == "sphere":
        verts = torch.from_numpy(meshes.lumpy_sphere()).to(device)
        uvs = None
    else:
        n = cfg["n"]
        if rank == 0:
            verts = meshes.random_triangles_torch(n, cfg["extent"], seed=1, device=device)
            uvs = meshes.random_uvs_torch(n, seed=2, device=device) if cfg.get("textured") else None
        else:
            verts = torch.empty((n, 9), dtype=torch.float32, device=device)
            uvs = torch.empty((n, 6), dtype=torch.float32, device=device) if cfg.get("textured") else None
        if distributed:
            slabs.broadcast_mesh([verts, uvs], src=0)
    n_tri = verts.shape[0]
    textures = []
    if uvs is not None:
        textures = [(torch.from_numpy(meshes.random_texture(256, 256, 3)).to(device), o2v.UV_WRAP)]

    bounds = slabs.equal_slabs(S, world)
    z0, z1 = slabs.my_slab(bounds, rank)
    params = o2v.make_params(resolution=cfg["resolution"], supersampling=cfg["supersampling"],
                             strategy=cfg["strategy"], bounds=cfg["bounds"], slab=(z0, z1) if distributed else None)

    def sync_all():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- device-resident steps ----
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stats = None
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # started before the warm-up (nvidia-smi takes a while to come up); samples are windowed
    for _ in range(args.warmup):
        stats = engine.voxelize_device(verts, params, uvs=uvs, textures=textures)
    sync_all()
    sampler.mark()
    kernel_ms, setup_ms, clip_ms, classify_ms, launches = [], [], [], [], 0
    start.record()
    for _ in range(args.steps):
        stats = engine.voxelize_device(verts, params, uvs=uvs, textures=textures)
        kernel_ms.append(stats["ms_voxelize"])
        clip_ms.append(stats["ms_clip"])
        classify_ms.append(stats["ms_classify"])
        setup_ms.append(stats["ms_setup"])
        launches += stats["kernel_launches"]
    stop.record()
    sync_all()
    elapsed_ms = start.elapsed_time(stop)
    clocks = sampler.stop() if rank == 0 else None
    occupancy_path = bool(stats["occupancy_path"])

    # The same workload with the occupancy-only path switched off (weights and colours folded for every voxel): what a
    # coloured / textured mesh of this shape costs.  Reported beside the headline, not as the headline.
    weighted = None
    if occupancy_path and not distributed:
        wparams = o2v.make_params(resolution=cfg["resolution"], supersampling=cfg["supersampling"],
                                  strategy=cfg["strategy"], bounds=cfg["bounds"], occupancy_path=0)
        for _ in range(2):
            wstats = engine.voxelize_device(verts, wparams, uvs=uvs, textures=textures)
        sync_all()
        wsteps = max(1, min(args.steps, 5))
        wstart, wstop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wstart.record()
        for _ in range(wsteps):
            wstats = engine.voxelize_device(verts, wparams, uvs=uvs, textures=textures)
        wstop.record()
        sync_all()
        wms = wstart.elapsed_time(wstop) / wsteps
        if wstats["voxels"] != stats["voxels"]:
            raise RuntimeError("weighted path emitted %d voxels, occupancy path %d" % (wstats["voxels"], stats["voxels"]))
        weighted = {"value": n_tri / (wms * 1e-3) / 1e6, "unit": "Mtri/s", "ms_per_step": wms,
                    "clip_calls": wstats["clip_calls"], "contributions": wstats["contributions"],
                    "ms_clip_kernel": wstats["ms_clip"],
                    "note": "same workload with occupancy_path=0: every (triangle, voxel) weight folded in reference "
                            "order (what coloured / textured meshes cost); identical output"}
        launches += 0  # not part of the timed region above

    counts = [stats["voxels"], stats["contributions"], stats["clip_calls"], stats["leaves"], launches]
    tile_split = (stats["light_tiles"], stats["heavy_tiles"])
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        counts = slabs.allreduce_counts(counts, device)
    elapsed_ms = float(t.item())
    voxels, contributions, clip_calls, leaves, launches = counts
    ms_per_step = elapsed_ms / args.steps

    # ---- end to end with HOST buffers (H2D + kernels + D2H inside the timed region) ----
    # N = 1: the reference-facing C API (obj2voxel instance, bulk input, voxel callback).
    # N > 1: every rank uploads its 1/N share of the host triangle array from pinned memory, the shares are all-gathered
    #        over NVLink (the "triangles broadcast once" of the slab scheme; no voxel data is exchanged), each rank
    #        voxelizes its Z-slab through the Engine API and downloads its own voxels into pinned host memory.
    pinned_verts = torch.empty(verts.shape, dtype=verts.dtype, pin_memory=True)
    pinned_verts.copy_(verts)
    host_verts = pinned_verts.numpy()
    host_uvs = None
    pinned_uvs = None
    if uvs is not None:
        pinned_uvs = torch.empty(uvs.shape, dtype=uvs.dtype, pin_memory=True)
        pinned_uvs.copy_(uvs)
        host_uvs = pinned_uvs.numpy()
    e2e_times, e2e_voxels = [], 0
    e2e_steps = max(1, min(args.steps, 5))
    if not distributed:
        tex_obj = o2v.Texture(meshes.random_texture(256, 256, 3), wrap=o2v.UV_WRAP) if uvs is not None else None
        lib = o2v.load()
        lib.obj2voxel_set_log_level(o2v._lib.LOG_ERROR)
        for step in range(1 + e2e_steps):
            inst = o2v.Instance()
            inst.set_input_triangles(host_verts, uvs=host_uvs, texture=tex_obj)
            received = {"n": 0}

            def on_voxels(_data, _quads, count, received=received):
                received["n"] += count
                return True

            cb = o2v._lib.VOXEL_CALLBACK(on_voxels)
            lib.obj2voxel_set_output_callback(inst.handle, cb, None)
            inst.set_resolution(cfg["resolution"])
            inst.set_supersampling(cfg["supersampling"])
            inst.set_color_strategy(cfg["strategy"])
            if cfg["bounds"] is not None:
                inst.set_mesh_boundaries(cfg["bounds"])
            sync_all()
            t0 = time.perf_counter()
            err = inst.voxelize()
            torch.cuda.synchronize(device)
            dt = time.perf_counter() - t0
            inst.free()
            if err != 0:
                raise RuntimeError("obj2voxel_voxelize failed with error %d" % err)
            if step > 0:
                e2e_times.append(dt)
                e2e_voxels = received["n"]
        e2e_api = ("obj2voxel_b200_set_input_triangles + obj2voxel_voxelize + voxel callback (the job runs in 4 z parts: "
                   "the download of one under the kernels of the next)")
    else:
        per_rank = -(-n_tri // world)
        lo, hi = min(rank * per_rank, n_tri), min((rank + 1) * per_rank, n_tri)
        share_v = torch.zeros((per_rank, 9), dtype=torch.float32, device=device)
        full_v = torch.empty((per_rank * world, 9), dtype=torch.float32, device=device)
        share_u = full_u = None
        if uvs is not None:
            share_u = torch.zeros((per_rank, 6), dtype=torch.float32, device=device)
            full_u = torch.empty((per_rank * world, 6), dtype=torch.float32, device=device)
        out_pinned = torch.empty((max(int(stats["voxels"] * 1.1) + 1024, 1), 4), dtype=torch.int32, pin_memory=True)
        out_np = out_pinned.numpy().view(np.uint32)
        stream = torch.cuda.current_stream(device).cuda_stream
        for step in range(1 + e2e_steps):
            sync_all()
            t0 = time.perf_counter()
            share_v[: hi - lo].copy_(pinned_verts[lo:hi], non_blocking=True)
            dist.all_gather_into_tensor(full_v, share_v)
            if uvs is not None:
                share_u[: hi - lo].copy_(pinned_uvs[lo:hi], non_blocking=True)
                dist.all_gather_into_tensor(full_u, share_u)
            st = engine.voxelize_device(full_v[:n_tri], params, uvs=None if uvs is None else full_u[:n_tri],
                                        textures=textures)
            got = engine.download(out=out_np, stream=stream)
            torch.cuda.synchronize(device)
            dt = time.perf_counter() - t0
            if step > 0:
                e2e_times.append(dt)
                e2e_voxels = len(got)
        e2e_api = ("pinned host share -> H2D -> all_gather (NVLink) -> Engine.voxelize_device(slab) -> "
                   "Engine.download to pinned host")
    e2e_t = torch.tensor([float(np.mean(e2e_times))], dtype=torch.float64, device=device)
    e2e_counts = [e2e_voxels]
    if distributed:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        e2e_counts = slabs.allreduce_counts(e2e_counts, device)
    e2e_seconds = float(e2e_t.item())
    if e2e_counts[0] != voxels:
        raise RuntimeError("e2e voxel count %d != device-resident count %d" % (e2e_counts[0], voxels))

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
        alg_bytes = 16 * stats["voxels"] + tri_bytes * n_tri
        achieved = alg_bytes / (d_ms * 1e-3) / 1e9 if d_ms > 0 else 0.0
        traffic = profiled_traffic(workload, dominant)
        baseline = cpu_baseline(cfg) if world == 1 else None
        line = {
            "metric": "triangles_per_second", "value": n_tri / (ms_per_step * 1e-3) / 1e6, "unit": "Mtri/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "description": WORKLOADS[workload], "triangles": n_tri,
                       "resolution": cfg["resolution"], "supersampling": cfg["supersampling"],
                       "strategy": "blend" if cfg["strategy"] else "max",
                       "partition": "z-slabs %s" % bounds if distributed else "single GPU, whole grid",
                       "l2": "inputs (%d MB of triangles) exceed the 126 MB L2; no explicit flush" %
                             (n_tri * tri_bytes // 1000000)},
            "mvoxel_per_s": voxels / (ms_per_step * 1e-3) / 1e6, "voxels": voxels, "contributions": contributions,
            "clip_calls": clip_calls, "leaves": leaves, "light_tiles_rank0": tile_split[0],
            "heavy_tiles_rank0": tile_split[1],
            "ms_setup_rank0": float(np.mean(setup_ms)), "ms_voxelize_rank0": k_ms, "ms_clip_kernel_rank0": c_ms,
            "ms_classify_kernel_rank0": f_ms,
            "path": "occupancy-only (every triangle MATERIALLESS: output colour is white whatever the weights)"
                    if occupancy_path else "weighted fold",
            "hbm_write_gbs": 16 * stats["voxels"] / (ms_per_step * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (traffic // world) if traffic else None,
                         "peak_source": peak_source,
                         "note": "algorithmic bytes = 16 B x voxels + %d B x triangles per launch / CUDA-event duration of "
                                 "the dominant kernel; the kernel is bound by instruction issue / latency of the "
                                 "triangle-box predicates, not by HBM — see DESIGN.md section 4" % tri_bytes},
            "e2e": {"value": n_tri / e2e_seconds / 1e6, "unit": "Mtri/s", "ms_per_step": e2e_seconds * 1e3,
                    "h2d_bytes_per_step": int(n_tri * tri_bytes), "d2h_bytes_per_step": int(16 * voxels),
                    "api": e2e_api},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if weighted is not None:
            line["weighted_path"] = weighted
        if baseline is not None:
            line["cpu_baseline"] = baseline
        print(json.dumps(line), flush=True)
    engine.close()
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    return 0


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
