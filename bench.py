#!/usr/bin/env python
"""Benchmark of the path-tracing hot path (BASELINE.json metric: Mrays/s and samples/s, % of the traversal roofline).

    python bench.py --gpus N --steps K --warmup W [--config c2] [--impl reference]

A *step* is one progressive pass (render_all_tiles, Raytracer/raytracer.cpp:692-757) of the named configuration:
every pixel of the frame x its samples-per-pixel through ray generation, TLAS/BLAS traversal, shading/NEE,
Russian roulette and the Mitchell-Netravali splat.  N=1 renders BASELINE config 2 (1,310,720-triangle displaced
icosphere, 1920x1080, 64 spp).  With N>1 (torchrun, one rank per GPU) the same frame is split into interleaved
64-row blocks over the ranks, the scene is replicated, and the partial films are summed with ONE NCCL reduce per
pass (strong scaling).

`value`     device-timed whole-job Mrays/s with the scene resident in HBM (CUDA events, max over ranks).
`e2e`       the same metric through the public C-ABI with host buffers: every step re-uploads the host scene
            (bpt_upload_scene: flatten + H2D), renders, and downloads the film (D2H) inside the timed region.
`roofline`  persistent_trace (all traversal launches): algorithmic bytes (reference binary-BVH visit counts x SURVEY 8d
            byte sizes) / their summed launch time, against the measured HBM bandwidth.
`cpu_baseline` / `--impl reference`: the reference's own tile-multithreaded CPU renderer (oracle/_ref, built from
            /root/reference unmodified) on this box's host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_NODE, BYTES_INSTANCE, BYTES_TRI = 32, 104, 40
BLOCK_ROWS = 8                                           # interleave granularity: 1080 rows = 135 blocks -> <= 1 % imbalance at 8 ranks (64-row tiles, raytracer.cpp:1661, give 17 blocks = 41 %)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def my_rows(h, rank, world):
    """interleaved BLOCK_ROWS-row blocks: block b belongs to rank b % world"""
    if world == 1:
        return [(0, h)]
    out = []
    for b, y0 in enumerate(range(0, h, BLOCK_ROWS)):
        if b % world == rank:
            out.append((y0, min(h, y0 + BLOCK_ROWS)))
    return out


def algorithmic_bytes(st, shadow):
    """SURVEY 8d B_ray summed over a stats snapshot, for the closest-hit kernel (shadow=False) or the shadow kernel."""
    d = st.as_dict()
    if shadow:
        tl, inst, bl, tri = (d["shadow_tlas_node_pops"], d["shadow_instances_visited"],
                             d["shadow_mesh_bvh_traversals"], d["shadow_triangles_tested"])
    else:
        tl = d["tlas_node_pops"] - d["shadow_tlas_node_pops"]
        inst = d["instances_visited"] - d["shadow_instances_visited"]
        bl = d["mesh_bvh_traversals"] - d["shadow_mesh_bvh_traversals"]
        tri = d["triangles_tested"] - d["shadow_triangles_tested"]
    return BYTES_NODE * tl + BYTES_INSTANCE * inst + BYTES_NODE * bl + BYTES_TRI * tri


def cpu_reference_run(cfg_key, w, h, seconds_target=15.0, threads=None, scale_div=1):
    """Time the reference's own WorkQueue renderer (verbatim, per-tile seeding) on a bounded sample of the workload."""
    from buas_pathtracer_b200 import scenes
    from oracle import ref_oracle
    cfg = scenes.CONFIGS[cfg_key]
    nproc = os.cpu_count() or 1
    if threads is None:
        threads = nproc + nproc // 4                       # raytracer.cpp:1580-1592
    ref = ref_oracle.RefScene()
    cfg["build"](ref, w, h)
    # calibrate with 1 spp on a reduced frame, then size spp for ~seconds_target of CPU work
    cw, ch = max(64, w // 4), max(36, h // 4)
    ref2 = ref_oracle.RefScene()
    cfg["build"](ref2, cw, ch)
    _, sec, _ = ref2.render_threaded(cw, ch, 1, threads, want_film=False)
    rate = cw * ch / max(sec, 1e-6)
    spp = int(max(1, min(cfg["spp"], round(rate * seconds_target / (w * h)))))
    _, sec, st = ref.render_threaded(w, h, spp, threads, want_film=False)
    samples = w * h * spp
    return {"seconds": sec, "samples": samples, "spp": spp, "threads": threads, "nproc": nproc,
            "samples_per_s": samples / sec}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from buas_pathtracer_b200 import scenes
    cfg = scenes.CONFIGS[args.config]
    w, h, spp = cfg["w"], cfg["h"], cfg["spp"]
    config = {"workload": cfg["name"], "key": args.config, "width": w, "height": h, "spp": spp,
              "integrator": "Advanced Pathtracer", "filter": "Mitchell Netravali", "sampler": "Stratified",
              "seeding": "per-pixel counter-based", "partition": f"interleaved {BLOCK_ROWS}-row blocks over {world} GPU(s)",
              "l2_policy": "working set (BVH+triangles+path state+film) exceeds L2; no flush needed"}

    # ------------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        vals = []
        info = None
        for i in range(args.warmup + args.steps):
            info = cpu_reference_run(args.config, w, h, seconds_target=8.0)
            if i >= args.warmup:
                vals.append(info)
        sec = float(np.mean([v["seconds"] for v in vals]))
        samples = vals[0]["samples"]
        rps = float(os.environ.get("BPT_RAYS_PER_SAMPLE", "0")) or None
        # rays/sample of this workload is a property of the integrator + scene; measured by the GPU arm's counters
        # (parity tests show identical per-sample ray counts) and cached next to the bench for the reference arm.
        cache = os.path.join(ROOT, "profiles", f"rays_per_sample_{args.config}.json")
        if rps is None and os.path.exists(cache):
            rps = json.load(open(cache))["rays_per_sample"]
        if rps is None:
            rps = 1.0
        value = samples * rps / sec / 1e6
        line = {"impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "samples_per_s": samples / sec, "rays_per_sample": rps,
                "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": info["threads"], "kind": "reference",
                                 "sample": f"{w}x{h} at {info['spp']} spp ({samples} samples) per step, verbatim WorkQueue "
                                           f"renderer, {info['threads']} worker threads on {info['nproc']} logical CPUs"},
                "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------------------------------------------------
    import torch
    import buas_pathtracer_b200 as B

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    scene = B.Scene()
    cfg["build"](scene, w, h)
    r = B.Renderer(local_rank)
    r.upload_scene(scene)
    # the film lives in a torch tensor so the NCCL reduce can run on it in place
    film = torch.zeros((h, w, 4), dtype=torch.float32, device=f"cuda:{local_rank}")
    r.film_use_external(film.data_ptr(), w, h)
    bands = my_rows(h, rank, world)

    def render(frame_count):
        r.render_pass_bands(spp, bands, frame_count=frame_count)   # this rank's row blocks as one workload

    def step(frame_count):
        render(frame_count)
        r.sync()
        if dist is not None:
            dist.reduce(film, dst=0, op=dist.ReduceOp.SUM)     # one collective per progressive pass

    # --- counting pass (untimed): reference-unit visit counts for the roofline, rays per pass ---
    r.stats_enable(True)
    r.get_stats(reset=True)
    render(0)
    r.sync()
    st_counts = r.get_stats(reset=True)
    r.stats_enable(False)
    film.zero_()

    for i in range(args.warmup):
        step(i * spp)
    film.zero_()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()

    # --- timed region: K steps, CUDA events on this rank, max over ranks ---
    r.get_stats(reset=True)
    clocks = ClockSampler(local_rank)
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    trace_ms = shadow_ms = shade_ms = splat_ms = raygen_ms = 0.0
    launches = trace_launches = 0
    for i in range(args.steps):
        step(0)                                       # same frame_count -> same rays as the counting pass
    ev1.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    clk = clocks.stop()
    ms_total = ev0.elapsed_time(ev1)
    st_timed = r.get_stats(reset=True)

    # per-kernel breakdown: one more pass, batches serialised so that each kernel's own duration is measured
    r.set_detailed_timing(True)                       # one batch at a time, CUDA events around every kernel
    render(0)
    r.sync()
    t = r.pass_timing()
    trace_ms, shadow_ms, shade_ms, splat_ms, raygen_ms = t.trace_ms, t.shadow_ms, t.shade_ms, t.splat_ms, t.raygen_ms
    launches, trace_launches = t.kernel_launches, t.trace_launches
    r.set_detailed_timing(False)

    rays_step = st_timed.rays / max(1, args.steps)
    samples_step = sum((y1 - y0) for y0, y1 in bands) * w * spp
    t_ms = torch.tensor([ms_total, float(rays_step), float(samples_step)], dtype=torch.float64, device=f"cuda:{local_rank}")
    if dist is not None:
        mx = t_ms.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t_ms.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total = float(mx[0]); rays_all = float(sm[1]); samples_all = float(sm[2])
    else:
        rays_all, samples_all = float(rays_step), float(samples_step)
    ms_per_step = ms_total / args.steps
    value = rays_all / (ms_per_step * 1e-3) / 1e6

    # --- end-to-end: host scene -> upload -> render -> film back on the host, every step ---
    e2e = None
    if not args.no_e2e:
        host_film = np.empty((h, w, 4), np.float32)
        r2 = r
        r2.transfer_bytes(reset=True)
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(max(1, min(args.steps, 3))):
            r2.upload_scene(scene)
            film.zero_()
            step(0)
            if rank == 0:
                r2.download_film(host_film)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        n_e2e = max(1, min(args.steps, 3))
        dt = (time.perf_counter() - t0) / n_e2e
        if dist is not None:
            tt = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local_rank}")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt[0])
        h2d, d2h = r2.transfer_bytes(reset=True)
        e2e = {"value": rays_all / dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": h2d // n_e2e,
               "d2h_bytes_per_step": d2h // n_e2e, "ms_per_step": dt * 1e3,
               "what": "per step: bpt_upload_scene (whole scene from page-locked host memory, re-laid-out on the device) + bpt_render_pass + bpt_download_film to a host array"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # --- roofline of the dominant kernel: persistent_trace (trace.cuh), launched as k_trace_closest (bounce 0),
    #     k_trace_merged (extension rays of bounce b + shadow rays of bounce b-1) and k_trace_shadow (last bounce).
    #     algorithmic bytes = reference binary-BVH visit counts of ALL rays x SURVEY 8d sizes; time = the summed
    #     CUDA-event duration of all those launches in one pass (events on the launching stream).
    peak, peak_src = measured_peaks()
    bytes_closest = algorithmic_bytes(st_counts, shadow=False)
    bytes_shadow = algorithmic_bytes(st_counts, shadow=True)
    closest_rays = st_counts.rays - st_counts.shadow_rays
    trav_ms = trace_ms + shadow_ms
    achieved = (bytes_closest + bytes_shadow) / (trav_ms * 1e-3) / 1e9 if trav_ms > 0 else None
    roofline = {"bound": "hbm", "kernel": "persistent_trace (all k_trace_closest / k_trace_shadow / k_trace_merged / k_tail launches of a pass)",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_ray": (bytes_closest + bytes_shadow) / max(1, st_counts.rays),
                "algorithmic_bytes_per_launch": (bytes_closest + bytes_shadow) / max(1, trace_launches),
                "closest_bytes_per_ray": bytes_closest / max(1, closest_rays),
                "shadow_bytes_per_ray": bytes_shadow / max(1, st_counts.shadow_rays),
                "kernel_ms_per_step": trav_ms, "launches_per_step": trace_launches,
                "avg_launch_ms": trav_ms / max(1, trace_launches),
                "stage_ms_per_step": {"raygen": raygen_ms, "trace_closest_merged_tail": trace_ms, "shade": shade_ms,
                                      "trace_shadow": shadow_ms, "splat": splat_ms}}
    traffic_file = os.path.join(ROOT, "profiles", "trace_dram_bytes.json")
    if os.path.exists(traffic_file) and args.config == "c2" and world == 1:     # measured for exactly this workload
        tf = json.load(open(traffic_file))
        roofline["traffic"] = tf.get("dram_bytes_per_launch")
        roofline["traffic_source"] = tf.get("source")

    rps = rays_all / samples_all
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    try:
        with open(os.path.join(ROOT, "profiles", f"rays_per_sample_{args.config}.json"), "w") as f:
            json.dump({"rays_per_sample": rps, "config": args.config}, f)
    except OSError:
        pass

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            info = cpu_reference_run(args.config, w, h, seconds_target=15.0)
            cpu_baseline = {"value": info["samples_per_s"] * rps / 1e6, "unit": "Mrays/s", "cores": info["threads"],
                            "kind": "reference", "samples_per_s": info["samples_per_s"],
                            "sample": f"{w}x{h} at {info['spp']} spp ({info['samples']} samples), the reference's verbatim "
                                      f"WorkQueue renderer (per-tile seeding), {info['threads']} worker threads on "
                                      f"{info['nproc']} logical CPUs; rays = samples x GPU-counted rays/sample"}
        except Exception as e:  # the oracle is a checker, never a dependency of the measured path
            cpu_baseline = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}

    line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "samples_per_s": samples_all / (ms_per_step * 1e-3), "rays_per_step": rays_all, "rays_per_sample": rps,
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches * args.steps), "roofline": roofline,
            "cpu_baseline": cpu_baseline}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
