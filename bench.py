#!/usr/bin/env python
"""Benchmark of the path-tracing hot path (BASELINE.json metric: Mrays/s and samples/s, % of the traversal roofline).

    python bench.py --gpus N --steps K --warmup W [--config c2] [--impl reference]

A *step* is one progressive pass (render_all_tiles, Raytracer/raytracer.cpp:692-757) of the named configuration:
every pixel of the frame x its samples-per-pixel through ray generation, TLAS/BLAS traversal, shading/NEE,
Russian roulette and the Mitchell-Netravali splat.  N=1 renders BASELINE config 2 (1,310,720-triangle displaced
icosphere, 1920x1080, 64 spp).  With N>1 (torchrun, one rank per GPU) the same frame is split into interleaved
8-row blocks over the ranks, the scene is replicated, and the partial films are summed with ONE NCCL reduce per
pass by the library itself (bpt_reduce_film, include/bpt.h section 3; strong scaling).
The K timed steps are enqueued back to back and waited for once (barrier + synchronize on both sides of the K steps):
the library overlaps consecutive passes on the device and keeps every film reader ordered.

`value`     device-timed whole-job Mrays/s with the scene resident in HBM (CUDA events, max over ranks).
`e2e`       the same metric through the public C-ABI with host buffers: every step re-uploads the host scene
            (bpt_upload_scene_async: flatten + H2D), renders, and reads the film back (bpt_download_film_async: D2H)
            inside the timed region; the copies of neighbouring steps overlap the rendering.
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

    def mark(self):
        """the timed region starts here: only samples taken from now on are reported (nvidia-smi needs ~0.2 s to come up, so it is
        started ahead of the warm-up)"""
        self.first = len(self.rows)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows[getattr(self, "first", 0):]:
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


def reference_scene(cfg_key, w, h):
    """The configuration's scene inside the reference's own Scene (oracle/_ref).  Inputs come from the standalone
    inputs library: nothing of the product (libbpt.so) is loaded on this path."""
    from buas_pathtracer_b200 import scenes
    from oracle import ref_oracle, ref_inputs
    scenes.INPUTS = ref_inputs
    ref = ref_oracle.RefScene()
    scenes.CONFIGS[cfg_key]["build"](ref, w, h)
    return ref


def reference_rays_per_sample(ref, w, h, spp, rows=16):
    """Rays per sample of the workload, counted by the REFERENCE itself (untimed): its single-threaded render with the
    oracle's intersect_scene / intersect_shadow_ray call counter, on `rows` evenly spaced full rows of the frame at the
    configuration's spp."""
    total = n = 0
    step = max(1, h // rows)
    for y in range(step // 2, h, step):
        _, rec = ref.render_parity(w, h, spp, rect=(0, y, w, y + 1), records=True)
        total += int(rec["rays"].astype(np.int64).sum())
        n += rec.shape[0]
    return total / max(1, n), n


def cpu_reference_run(cfg_key, w, h, seconds_target=15.0, threads=None, ref=None, spp=None):
    """Time the reference's own WorkQueue renderer (verbatim, per-tile seeding) on a bounded sample of the workload."""
    from buas_pathtracer_b200 import scenes
    cfg = scenes.CONFIGS[cfg_key]
    nproc = os.cpu_count() or 1
    if threads is None:
        threads = nproc + nproc // 4                       # raytracer.cpp:1580-1592
    if ref is None:
        ref = reference_scene(cfg_key, w, h)
    if spp is None:
        # calibrate with 1 spp on a reduced frame, then size spp for ~seconds_target of CPU work
        cw, ch = max(64, w // 4), max(36, h // 4)
        ref2 = reference_scene(cfg_key, cw, ch)
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
    ap.add_argument("--steps", type=int, default=10)
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
        # The reference's own CPU renderer (oracle/_ref = the reference's translation units compiled here, unmodified)
        # through its own WorkQueue, all host threads, on this arm's config.  A step = one pass over the full frame at a
        # bounded spp (stated in config["spp_per_step_run"]; samples/s of this renderer does not depend on spp).
        # Nothing of the product is loaded: inputs come from oracle/_ref/libbpt_inputs.so, rays are counted by the
        # reference's own call counter in an untimed pass.
        if rank != 0:
            return 0
        ref = reference_scene(args.config, w, h)
        rps, rps_n = reference_rays_per_sample(ref, w, h, spp)
        vals = []
        info = None
        spp_run = None
        for i in range(args.warmup + args.steps):
            info = cpu_reference_run(args.config, w, h, seconds_target=8.0, ref=ref, spp=spp_run)
            spp_run = info["spp"]                      # calibrated once, then every step runs the same sample
            if i >= args.warmup:
                vals.append(info)
        sec = float(np.mean([v["seconds"] for v in vals]))
        samples = vals[0]["samples"]
        value = samples * rps / sec / 1e6
        config = dict(config, spp_per_step_run=info["spp"], l2_policy="n/a (CPU renderer)",
                      partition="reference WorkQueue: 64x64 tiles over host threads (raytracer.cpp:551-757)",
                      seeding="per-tile stream (raytracer.cpp:588-591)")
        line = {"impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "samples_per_s": samples / sec, "rays_per_sample": rps,
                "rays_per_sample_source": f"counted by the reference (oracle call counter) on {rps_n} samples of this frame, untimed",
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
    r.film_resize(w, h)
    bands = my_rows(h, rank, world)

    # Multi-GPU lives in the library (include/bpt.h section 3): every rank accumulates its row blocks into its own
    # film, bpt_reduce_film sums the films on rank 0 with ONE ncclReduce per progressive pass, enqueued on the
    # library's stream right behind the pass (no host synchronisation in between).  torch.distributed only carries the
    # NCCL id to the ranks, the barriers and the max-over-ranks of the timings.
    comm = None
    if dist is not None:
        uid = torch.from_numpy(B.Renderer.nccl_unique_id() if rank == 0 else np.zeros(128, np.uint8)).to(f"cuda:{local_rank}")
        dist.broadcast(uid, src=0)
        comm = r.nccl_comm_init_rank(uid.cpu().numpy(), world, rank)

    def render(frame_count, n_spp=spp):
        r.render_pass_bands(n_spp, bands, frame_count=frame_count)   # this rank's row blocks as one workload

    def step(frame_count):
        # One progressive pass (+ its reduce), enqueued; the host does not wait per step, like the reference's render loop
        # that starts the next pass as soon as the previous one is handed to the display (raytracer.cpp:692-757).  The
        # library orders what must be ordered on the device: pass k+1's splats wait for pass k's reduce / film readers.
        render(frame_count)
        if comm is not None:
            r.reduce_film(comm, 0)                         # one collective per progressive pass, behind the pass on its stream

    # --- multi-GPU correctness on the hardware (untimed): the N-rank reduced film against rank 0 rendering the whole
    #     frame alone, same seeds; they differ only in the order of float additions (atomics + the reduce)
    multi_gpu_check = None
    if comm is not None:
        chk_spp = 2
        render(0, chk_spp)
        r.reduce_film(comm, 0)
        r.sync()
        dist.barrier()
        if rank == 0:
            reduced = r.download_reduced_film()
            r.film_clear()
            r.render_pass(chk_spp, frame_count=0)
            alone = r.download_film()
            err = np.abs(reduced.astype(np.float64) - alone.astype(np.float64))
            scale = np.abs(alone.astype(np.float64)) + 1e-3 * float(np.mean(np.abs(alone)))
            multi_gpu_check = {"what": f"{world}-rank reduced film vs rank 0 rendering the full frame alone, {chk_spp} spp",
                               "max_rel_err": float(np.max(err / scale)),
                               "ok": bool(np.allclose(reduced, alone, rtol=1e-4, atol=1e-5))}
        dist.barrier()
        r.film_clear()
        r.sync()

    # --- counting pass (untimed): reference-unit visit counts for the roofline, rays per pass ---
    r.stats_enable(True)
    r.get_stats(reset=True)
    render(0)
    r.sync()
    settled_rays, settled_bytes = r.ray_prefilter_stats()      # shadow rays a timed pass settles inside k_shade (they never reach a traversal launch)
    st_counts = r.get_stats(reset=True)
    r.stats_enable(False)
    r.film_clear()

    clocks = ClockSampler(local_rank)
    clocks.start()
    for i in range(args.warmup):                      # enqueued back to back like the timed steps (same batch shapes, path state allocated)
        step(i * spp)
    r.sync()
    r.film_clear()
    r.sync()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()

    # --- timed region: K steps, CUDA events on this rank, max over ranks ---
    r.get_stats(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    clocks.mark()
    ev0.record()
    trace_ms = shadow_ms = shade_ms = splat_ms = raygen_ms = 0.0
    launches = trace_launches = 0
    for i in range(args.steps):
        step(0)                                       # same frame_count -> same rays as the counting pass
    r.sync()                                          # the library renders on its own streams: wait for all K steps, then stamp
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

    # --- end-to-end: host scene -> upload -> render -> film back on the host, every step -------------------------
    # The public calls a host application makes per frame: submit the scene (bpt_upload_scene_async: H2D from the
    # page-locked host scene into the scene buffer that is not being rendered), clear + render the pass (+ the NCCL
    # reduce), read the film back (bpt_download_film_async: snapshot to the front buffer, D2H to a page-locked host
    # array).  Every step moves all of its bytes inside the timed region; the copies of step i+1 / i-1 overlap the
    # rendering of step i on the copy streams, as in any double-buffered frame loop.  Timed by wall clock around the
    # whole loop incl. the final wait for the last film.
    e2e = None
    if not args.no_e2e:
        host_films = [np.empty((h, w, 4), np.float32) for _ in range(2)]
        if rank == 0:
            for hf in host_films:
                r.host_register(hf)
        n_e2e = max(3, args.steps)

        def frame_loop(n_frames):
            for i in range(n_frames):
                r.upload_scene_async(scene)
                r.film_clear()
                render(0)
                if comm is not None:
                    r.reduce_film(comm, 0)
                if rank == 0:
                    r.download_film_async(host_films[i & 1], reduced=comm is not None)
            if rank == 0:
                r.wait_download()
            r.sync()

        frame_loop(2)                                      # untimed warm-up: the second scene buffer and the front buffer get allocated
        r.transfer_bytes(reset=True)
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        frame_loop(n_e2e)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        if dist is not None:
            tt = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local_rank}")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt[0])
        h2d, d2h = r.transfer_bytes(reset=True)
        ok = bool(np.all(np.isfinite(host_films[(n_e2e - 1) & 1])) and float(host_films[(n_e2e - 1) & 1][..., 3].min()) > 0.0) if rank == 0 else True
        if rank == 0:
            for hf in host_films:
                r.host_unregister(hf)
        e2e = {"value": rays_all / dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": h2d // n_e2e,
               "d2h_bytes_per_step": d2h // n_e2e, "ms_per_step": dt * 1e3, "steps": n_e2e, "film_ok": ok,
               "what": "per step: bpt_upload_scene_async (whole scene from page-locked host memory into the inactive scene buffer, "
                       "re-laid-out on the device) + bpt_film_clear + bpt_render_pass_bands (+ bpt_reduce_film) + "
                       "bpt_download_film_async to a page-locked host array; copies of neighbouring steps overlap the rendering "
                       "(two scene buffers, front-buffer snapshot); wall clock over the loop incl. the wait for the last film"}

    # --- extra key at 8 GPUs: BASELINE config 5 (3840x2160, 1024 spp, the instanced scene), 2 timed passes -----------
    c5 = None
    if world >= 8 and args.config == "c2" and os.environ.get("BPT_BENCH_C5", "1") != "0":
        c5cfg = scenes.CONFIGS["c5"]
        w5, h5, spp5 = c5cfg["w"], c5cfg["h"], c5cfg["spp"]
        scene5 = B.Scene()
        c5cfg["build"](scene5, w5, h5)
        r.upload_scene(scene5)
        r.film_resize(w5, h5)
        bands5 = my_rows(h5, rank, world)

        def step5():
            r.render_pass_bands(spp5, bands5, frame_count=0)
            r.reduce_film(comm, 0)
            r.sync()

        step5()                                            # warm-up (path-state allocation for the larger batches)
        r.get_stats(reset=True)
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        n5 = 2
        for _ in range(n5):
            step5()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        st5 = r.get_stats(reset=True)
        v = torch.tensor([e0.elapsed_time(e1), float(st5.rays) / n5, float(sum(y1 - y0 for y0, y1 in bands5)) * w5 * spp5],
                         dtype=torch.float64, device=f"cuda:{local_rank}")
        mx = v.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = v.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms5 = float(mx[0]) / n5
        c5 = {"workload": c5cfg["name"], "n_gpus": world, "steps": n5, "ms_per_step": ms5,
              "value": float(sm[1]) / (ms5 * 1e-3) / 1e6, "unit": "Mrays/s",
              "samples_per_s": float(sm[2]) / (ms5 * 1e-3), "includes": "one NCCL reduce of the 132.7 MB film per pass"}

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
    bytes_shadow = algorithmic_bytes(st_counts, shadow=True) - settled_bytes     # what the traversal LAUNCHES process: the rays k_shade settles are not theirs
    closest_rays = st_counts.rays - st_counts.shadow_rays
    trav_ms = trace_ms + shadow_ms
    achieved = (bytes_closest + bytes_shadow) / (trav_ms * 1e-3) / 1e9 if trav_ms > 0 else None
    roofline = {"bound": "hbm", "kernel": "persistent_trace (all k_trace_closest / k_trace_shadow / k_trace_merged / k_tail launches of a pass)",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_ray": (bytes_closest + bytes_shadow) / max(1, st_counts.rays - settled_rays),
                "algorithmic_bytes_per_launch": (bytes_closest + bytes_shadow) / max(1, trace_launches),
                "closest_bytes_per_ray": bytes_closest / max(1, closest_rays),
                "shadow_bytes_per_ray": bytes_shadow / max(1, st_counts.shadow_rays - settled_rays),
                "shadow_rays_settled_in_k_shade": {"rays_per_step": settled_rays, "algorithmic_bytes_excluded": settled_bytes,
                                                   "what": "NEE shadow rays that never reach a mesh BLAS (plane hit / TLAS root missed / one-leaf TLAS: every item missed or a "
                                                           "sphere / box hit) are decided inside k_shade; their visits are not counted as traversal-kernel bytes"},
                "kernel_ms_per_step": trav_ms, "launches_per_step": trace_launches,
                "avg_launch_ms": trav_ms / max(1, trace_launches),
                "stage_ms_per_step": {"raygen": raygen_ms, "trace_closest_merged_tail": trace_ms, "shade": shade_ms,
                                      "trace_shadow": shadow_ms, "splat": splat_ms}}
    # traversal steps per ray in the reference's own units (node pops incl. failed box tests, triangles tested), from
    # the untimed counting pass
    dc = st_counts.as_dict()
    roofline["traversal_steps_per_ray"] = {
        "node_pops": (dc["tlas_node_pops"] + dc["mesh_bvh_traversals"]) / max(1, st_counts.rays),
        "inner_nodes_entered": dc["mesh_node_traversals"] / max(1, st_counts.rays),
        "leaves_visited": dc["mesh_leaf_traversals"] / max(1, st_counts.rays),
        "triangles_tested": dc["triangles_tested"] / max(1, st_counts.rays),
        "instances_visited": dc["instances_visited"] / max(1, st_counts.rays)}
    if roofline["frac"] is not None and roofline["frac"] > 1.2:
        roofline["note"] = ("algorithmic bytes are served from L1/L2 on this workload (the whole acceleration structure fits on chip): "
                            "the HBM convention of SURVEY 8d does not bound it; the traversal kernel is issue-bound (DESIGN.md 4.1)")
    traffic_file = os.path.join(ROOT, "profiles", "trace_dram_bytes.json")
    if os.path.exists(traffic_file) and args.config == "c2" and world == 1:     # measured for exactly this workload
        tf = json.load(open(traffic_file))
        # per launch like `achieved`: the capture's bytes per PASS over this run's launches per pass (the capture ran with a host wait per
        # pass, which splits a pass over both batch streams: twice the launches, the same rays)
        roofline["traffic"] = tf.get("dram_bytes_per_pass", 0) / max(1, trace_launches)
        roofline["traffic_per_pass"] = tf.get("dram_bytes_per_pass")
        roofline["traffic_source"] = tf.get("source")
        if achieved and roofline["traffic"] and roofline["traffic"] < 0.5*roofline["algorithmic_bytes_per_launch"]:
            roofline["note"] = ("`achieved` counts ALGORITHMIC bytes (SURVEY 8d convention); the measured DRAM traffic of these launches is "
                                f"{roofline['traffic']/roofline['algorithmic_bytes_per_launch']:.0%} of them -- the acceleration structure is served from L2 -- so the "
                                "fraction can exceed 1 and the kernel's actual bound is issue slots (issue_active_pct, active_lanes_of_32; DESIGN.md 4.1)")
        roofline["l2_bytes_per_launch"] = tf.get("l2_bytes_per_pass", 0) / max(1, trace_launches)
        mfile = os.path.join(ROOT, "profiles", "r2_trace_metrics.json")
        if os.path.exists(mfile):                                              # from the committed ncu capture of this workload
            tm = json.load(open(mfile))
            roofline["active_lanes_of_32"] = tm.get("active_lanes_of_32_weighted")
            roofline["issue_active_pct"] = tm.get("issue_active_pct_weighted")
            roofline["ncu_source"] = tm.get("source")

    rps = rays_all / samples_all

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            ref = reference_scene(args.config, w, h)
            ref_rps, ref_rps_n = reference_rays_per_sample(ref, w, h, spp)
            info = cpu_reference_run(args.config, w, h, seconds_target=15.0, ref=ref)
            cpu_baseline = {"value": info["samples_per_s"] * ref_rps / 1e6, "unit": "Mrays/s", "cores": info["threads"],
                            "kind": "reference", "samples_per_s": info["samples_per_s"], "rays_per_sample": ref_rps,
                            "sample": f"{w}x{h} at {info['spp']} spp ({info['samples']} samples), the reference's verbatim "
                                      f"WorkQueue renderer (per-tile seeding), {info['threads']} worker threads on "
                                      f"{info['nproc']} logical CPUs; rays = samples x rays/sample counted by the reference "
                                      f"itself on {ref_rps_n} samples of this frame"}
        except Exception as e:  # the oracle is a checker, never a dependency of the measured path
            cpu_baseline = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}

    line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "samples_per_s": samples_all / (ms_per_step * 1e-3), "rays_per_step": rays_all, "rays_per_sample": rps,
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches * args.steps), "roofline": roofline,
            "cpu_baseline": cpu_baseline}
    if multi_gpu_check is not None:
        line["multi_gpu_check"] = multi_gpu_check
    if c5 is not None:
        line["c5_at_8_gpus"] = c5
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
