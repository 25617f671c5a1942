#!/usr/bin/env python
"""bench.py — BASELINE.json metric "Mrays/sec and ms/spp at 1080p (device-timed) at 1/2/4/8 B200" on config 3:
the ~264 K-triangle procedural arcade mesh, 64 Lambert materials + 200 emissive triangles, (R)PathTracerSpectral
(hero wavelengths, HyperbolicPBRT, ACES_CG LUT), WithNEEAndMIS, rrRange [3, 8], 1920x1080, 1024 spp.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--spp S]

One STEP = one complete render of the configuration (1024 samples of every pixel, traced to completion, film
resolved on rank 0). `--gpus N` (one process per GPU under torchrun) is STRONG scaling: the BVH is built once per
rank, rank g renders the sample range [g S / N, (g + 1) S / N) of every pixel, and the per-GPU films are summed with one
NCCL all-reduce INSIDE the timed region. Random numbers are a function of (seed, pixel, sample index), so every N
renders the same image.

Prints ONE JSON line (rank 0):
  value      Mrays/s (closest-hit + shadow rays actually cast, all ranks) device-timed with CUDA events on the
             launching stream, max over ranks; `config.ms_per_spp_1080p` is the same time per sample.
  e2e        the same render through the reference-facing plugin: libTracerDLL_B200.so driven through TracerI (scene
             upload, CommitSurfaces, StartRender, DoRenderWork x burst passes, every film hand-off to pinned host memory,
             host-side accumulation), wall clock.
  roofline   the dominant kernel (wide closest-hit traversal inside the path tracer): SURVEY.md §8(d) bytes per ray x
             rays per launch / its mean launch time (sampled CUDA-event pairs inside the library) vs measured HBM peak.
  cpu_baseline / --impl reference: the UNMODIFIED reference path tracer (CPU device backend, oracle/_ref, all host
             threads) through the same TracerI driver on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1920, 1080
SPP = 1024
RR = (3, 8)
SAMPLE_MODE = "WithNEEAndMIS"
RENDERER = "PathTracerSpectral"
BURST = 32                     # PathTracerRendererBase::BurstSize (Tracer/PathTracerRendererBase.h:L163)
BYTES_CLOSEST = 764            # 64 B stream I/O + ceil(log2 N) = 18 x 36 B descent + 52 B leaf
BYTES_ANY = 736
BYTES_BOUNCE = 2030            # 528 B path state + one closest-hit + one any-hit traversal per path-bounce
BYTES_BUILD = 408              # SURVEY.md §8(d), per triangle
METRIC = "Mrays/sec and ms/spp at 1080p (device-timed)"
WORKLOAD = ("config 3: procedural arcade mesh 264038 tris, 64 Lambert + 200 emissive tris, (R)PathTracerSpectral "
            "WithNEEAndMIS rr[3,8], 1920x1080, 1024 spp; 1 step = 1 full render")
# Rays per camera path of this workload, measured by our renderer (identical estimator: images match the reference's):
# the reference casts one closest-hit ray per bounce and one shadow ray per NEE light sample (zero-valued ones too).
# Used only to express the reference's paths/s as rays/s; our own arm counts the rays it really casts and prints its
# measured ratios as config.rays_per_path for comparison.
REF_CLOSEST_PER_PATH = 3.13
REF_NEE_PER_PATH = 3.12


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def shard_samples(total, world, rank):
    base, extra = divmod(total, world)
    b = rank * base + min(rank, extra)
    return b, b + base + (1 if rank < extra else 0)


def build_scene():
    from mray_b200 import scenes
    p, i = scenes.arcade_mesh()
    pidx, pranges, pkeys, palb, prad, tri_mat = scenes.arcade_materials(p, i)
    return dict(p=p, i=i, pidx=pidx, pranges=pranges, pkeys=pkeys, palb=palb, prad=prad, tri_mat=tri_mat)


def driver_scene(sc):
    """The scene in the shape the TracerI driver uploads (one primitive batch per material)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    mat = np.where(sc["tri_mat"] < 0, len(sc["palb"]), sc["tri_mat"]).astype(np.uint32)
    bsc = O.batched_scene(sc["p"], sc["pidx"], mat)
    alb = np.concatenate([sc["palb"], np.zeros((1, 3), np.float32)])
    return O, bsc, alb


class ClockSampler:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        out, _ = self.proc.communicate()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import mray_b200
    from mray_b200 import capi, scenes, spectral

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL writes its version / debug lines to stdout by default; stdout carries exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        # ... and its "NCCL version" banner goes to fd 1 whatever that says: park stdout on stderr until the communicator exists
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.all_reduce(torch.zeros(1024, device="cuda")); torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved_stdout, 1); os.close(saved_stdout)
    if not spectral.available():
        raise SystemExit("bench.py: mray_b200/data/ACES_CG.mrspectra is missing (run __graft_entry__.build() where /root/reference exists)")
    ctx = mray_b200.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream)
    spp = args.spp

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    sc = build_scene()
    p, pidx = sc["p"], sc["pidx"]
    # ---- scene on the device: BVH build (device-timed inside the library; every rank builds its replica once) ----
    dp, di = torch.from_numpy(p).cuda(), torch.from_numpy(pidx.view(np.int32)).cuda()
    build_ms, acc = [], None
    for _ in range(3):
        if acc is not None:
            acc.close()
        acc = mray_b200.Accelerator(ctx, dp, di, prim_ranges=sc["pranges"], light_or_mat_keys=sc["pkeys"])
        build_ms.append(float(acc.info.buildMs))
    build_best = min(build_ms)
    spec = mray_b200.Spectrum(ctx, spectral.load(), "HyperbolicPBRT")
    s0, s1 = shard_samples(spp, world, rank)
    r = mray_b200.Renderer(ctx, acc, p.shape[0], pidx.shape[0], sc["palb"], sc["prad"], scenes.ARCADE_CAMERA, W, H, max(s1 - s0, 1),
                           sample_mode=SAMPLE_MODE, rr_range=RR, seed=0, spectrum=spec, sample_offset=s0, job_spp=spp,
                           film_filter="Gaussian", film_filter_radius=1.0)
    r.set_spp_limit(0)                               # passes are begun per step
    film = torch.zeros((4, H, W), dtype=torch.float32, device="cuda")
    if world > 1:                                    # NCCL builds its communicator on the first collective
        dist.all_reduce(torch.zeros(1024, device="cuda")); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    last = {"st": None}

    def step():
        ev[0].record(stream)
        r.begin_pass((0, 0), (W, H), s0, s1 - s0)
        st = r.run_pass(8)
        ctx.check(ctx.lib.mrb_renderer_read_film(ctx.handle, r.handle, film.data_ptr(), capi.MRB_MEM_DEVICE, 1))
        ev[1].record(stream)
        if world > 1:
            dist.all_reduce(film, op=dist.ReduceOp.SUM)   # NCCL over NVLink: the only data-path collective
        ev[2].record(stream)
        torch.cuda.synchronize()
        last["st"] = st
        return ev[0].elapsed_time(ev[2]), ev[1].elapsed_time(ev[2])

    warm = max(3, args.warmup)
    for _ in range(warm):
        step()
    barrier()
    base = last["st"]
    ctx.set_profiling(True, 16); ctx.get_profile()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    total_ms = reduce_ms = 0.0
    prof = None
    for _ in range(args.steps):
        a, b = step()
        total_ms += a; reduce_ms += b
        prof = ctx.get_profile()        # harvest the sampled event pairs (the pool holds 64 iterations' worth)
    barrier()
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    ctx.set_profiling(False)
    st = last["st"]
    fallback = ctx.last_fallback_stats
    closest = st.closestRays - base.closestRays
    shadow = st.shadowRays - base.shadowRays
    nee = st.neeSamples - base.neeSamples
    paths = st.pathsCompleted - base.pathsCompleted
    iters = st.iterations - base.iterations
    total_ms_max = max_over_ranks(total_ms)
    rays_all = sum_over_ranks(float(closest + shadow))
    paths_all = sum_over_ranks(float(paths))
    launches_all = int(sum_over_ranks(float(launches)))
    value = rays_all / (total_ms_max * 1e-3) / 1e6
    ms_per_step = total_ms_max / args.steps
    # rank 0 checks the reduced film: every pixel carries spp unit filter weights
    wmin, wmax = float(film[3].min().item()), float(film[3].max().item())

    # ---- roofline of the dominant kernel: wide closest-hit traversal inside the path tracer ----
    peak, peak_src = measured_peak()
    tc_ms, tc_n = prof["trace_closest"]
    rays_per_launch = closest / max(iters, 1)
    mean_launch_ms = tc_ms / max(tc_n, 1)
    achieved = rays_per_launch * BYTES_CLOSEST / (mean_launch_ms * 1e-3) / 1e9 if tc_n else 0.0
    per_iter = {k: (v[0] / v[1] if v[1] else 0.0) for k, v in prof.items()}
    iter_ms = sum(per_iter.values())
    shares = {k: round(v / iter_ms, 4) for k, v in per_iter.items()} if iter_ms > 0 else {}
    bounce_achieved = (closest * BYTES_BOUNCE) / (total_ms * 1e-3) / 1e9   # this rank: path-bounces = closest-hit rays
    r.close()

    # ---- config 2 sub-leg (traversal isolation on the same mesh): primary closest hit + AO closest / any hit ----
    sub = config2_leg(ctx, acc, sc, stream, rank, peak) if os.environ.get("MRB_BENCH_SKIP_CONFIG2") is None else None

    # ---- e2e: the same render through the plugin (TracerI), host buffers, every copy inside the timed region ----
    e2e = None
    if os.environ.get("MRB_BENCH_SKIP_E2E") is None:
        e2e = e2e_plugin(sc, spp, rank, world, local, rays_all / args.steps, barrier, max_over_ranks, args)

    out = {
        "metric": METRIC, "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": WORKLOAD if spp == SPP else WORKLOAD.replace("1024 spp", "%d spp" % spp),
            "ms_per_spp_1080p": round(ms_per_step / spp, 4),
            "mpaths_s": round(paths_all / (total_ms_max * 1e-3) / 1e6, 1),
            "rays_per_path": {"closest": round(closest / max(paths, 1), 3), "shadow_cast": round(shadow / max(paths, 1), 3),
                              "nee_samples": round(nee / max(paths, 1), 3)},
            "film_allreduce_ms_per_step": round(max_over_ranks(reduce_ms) / args.steps, 3),
            "iterations_per_step": int(iters // args.steps),
            "parallelism": ("sample ranges per rank (%d spp each), BVH replicated, NCCL all-reduce of the 33 MB film per step"
                            % (s1 - s0)) if world > 1 else "single GPU",
            "l2": "path state 2 073 600 slots x 212 B = 440 MB per iteration, far above the 126 MB L2",
            "film_weight_min_max": [wmin, wmax],
            "bvh_build_ms": round(build_best, 4), "bvh_build_mtris_s": round(pidx.shape[0] / (build_best * 1e-3) / 1e6, 1),
            "bvh_build_roofline_frac": round(pidx.shape[0] * BYTES_BUILD / (build_best * 1e-3) / 1e9 / peak, 4),
            "wide_nodes": int(acc.info.wideNodeCount), "exact_fallback_rays_last_cast": list(fallback),
            "kernel_ms_per_iteration": {k: round(v, 4) for k, v in per_iter.items()},
            "kernel_share_of_iteration": shares,
            "roofline_path_bounce": {"bound": "hbm", "achieved": round(bounce_achieved, 1), "peak": peak, "unit": "GB/s",
                                     "frac": round(bounce_achieved / peak, 4), "bytes_per_path_bounce": BYTES_BOUNCE},
            "config2_traversal": sub,
        },
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": launches_all,
        "roofline": {"bound": "hbm", "kernel": "KTraceWide<closest> (inside the path tracer)", "achieved": round(achieved, 1),
                     "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     # dram__bytes_read.sum + dram__bytes_write.sum of one mid-render launch, ncu --set full (profiles/)
                     "traffic": TRAFFIC_CLOSEST, "algorithmic_bytes_per_launch": round(rays_per_launch * BYTES_CLOSEST),
                     "rays_per_launch": round(rays_per_launch), "mean_launch_ms": round(mean_launch_ms, 4),
                     "launch_samples": tc_n, "bytes_per_ray": BYTES_CLOSEST, "peak_source": peak_src,
                     # second roofline (the kernel is issue-bound, not byte-bound): warp-instructions per second against
                     # 148 SMs x 4 schedulers x SM clock, one warp-instruction per scheduler per cycle
                     "issue": issue_roofline(rays_per_launch, mean_launch_ms, clocks)},
    }
    if rank == 0:
        if world == 1 and os.environ.get("MRB_BENCH_SKIP_CPU") is None:
            out["cpu_baseline"] = reference_pt_sample(sc, budget_s=20.0)
        print(json.dumps(out), flush=True)
    spec.close()
    acc.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


# From the ncu --set full capture of one mid-render KTraceWide<closest> launch (profiles/r2_pt_kernels.md, launch 0):
# dram__bytes_read.sum 93.0 MB + dram__bytes_write.sum 69.6 MB (rays in, hits out: the BVH itself stays in L2), and
# smsp__inst_executed.sum 774.2 M warp-instructions for the 2.06 M rays of that launch.
TRAFFIC_CLOSEST = 93_007_872 + 69_567_744
WARP_INST_PER_CLOSEST_RAY = 774_216_416 / 2_064_000


def issue_roofline(rays_per_launch, mean_launch_ms, clocks):
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    peak = 148 * 4 * sm_mhz * 1e6 / 1e9                              # G warp-instructions / s
    achieved = WARP_INST_PER_CLOSEST_RAY * rays_per_launch / (mean_launch_ms * 1e-3) / 1e9
    return {"bound": "issue", "achieved": round(achieved, 1), "peak": round(peak, 1), "unit": "Gwarp-inst/s",
            "frac": round(achieved / peak, 4), "warp_inst_per_ray": round(WARP_INST_PER_CLOSEST_RAY, 1),
            "lanes_per_inst": 20.0, "source": "profiles/r2_pt_kernels.md (ncu smsp__inst_executed.sum, "
            "smsp__thread_inst_executed_per_inst_executed.ratio)"}


def config2_leg(ctx, acc, sc, stream, rank, peak):
    """BASELINE config 2 on the same mesh: 1920x1080 primary closest hit + AO closest hit + AO any hit, 5 timed steps."""
    import torch
    from mray_b200 import capi, scenes
    p, pidx = sc["p"], sc["pidx"]
    rays_np = scenes.pinhole_rays(W, H, **scenes.ARCADE_CAMERA)
    n = rays_np.shape[0]
    d_primary = torch.from_numpy(rays_np).cuda()
    keys = torch.full((n, 4), -1, dtype=torch.int32, device="cuda"); hits = torch.zeros((n, 2), dtype=torch.float32, device="cuda")
    work = d_primary.clone()
    acc.cast_rays(keys, hits, work, None, capi.MRB_TRACE_WIDE)
    torch.cuda.synchronize()
    prim = keys.cpu().numpy().view(np.uint32)[:, 0] & 0x0FFFFFFF
    # prim keys index the accelerator's (material-sorted) index list
    e = acc.export_lbvh()
    diam = float(np.linalg.norm(e["accel_aabb"][3:] - e["accel_aabb"][:3]))
    ao_np = scenes.ao_rays(rays_np, np.where(keys.cpu().numpy().view(np.uint32)[:, 0] == 0xFFFFFFFF, 0xFFFFFFFF, prim).astype(np.uint32),
                           work.cpu().numpy()[:, 7], p, pidx, 0.15 * diam, seed_offset=rank * n)
    d_ao = torch.from_numpy(ao_np).cuda()
    words = (n + 31) // 32
    work_p, work_a = d_primary.clone(), d_ao.clone()
    keys_a = torch.full((n, 4), -1, dtype=torch.int32, device="cuda"); hits_a = torch.zeros((n, 2), dtype=torch.float32, device="cuda")
    bits = torch.full((words,), -1, dtype=torch.int32, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t = [0.0, 0.0, 0.0]
    steps = 5
    for k in range(3 + steps):
        work_p.copy_(d_primary); work_a.copy_(d_ao); keys.fill_(-1); keys_a.fill_(-1); bits.fill_(-1)
        ev[0].record(stream)
        acc.cast_rays(keys, hits, work_p, None, capi.MRB_TRACE_WIDE)
        ev[1].record(stream)
        acc.cast_rays(keys_a, hits_a, work_a, None, capi.MRB_TRACE_WIDE)
        ev[2].record(stream)
        acc.cast_visibility_rays(bits, d_ao, None, capi.MRB_TRACE_WIDE)
        ev[3].record(stream)
        torch.cuda.synchronize()
        if k >= 3:
            for q in range(3):
                t[q] += ev[q].elapsed_time(ev[q + 1])
    closest_gbs = 2 * n * BYTES_CLOSEST * steps / ((t[0] + t[1]) * 1e-3) / 1e9
    return {"workload": "config 2: %dx%d primary closest-hit + AO closest-hit + AO any-hit, %d rays per step" % (W, H, 3 * n),
            "mrays_s": round(3 * n * steps / (sum(t) * 1e-3) / 1e6, 1),
            "mrays_primary": round(n * steps / (t[0] * 1e-3) / 1e6, 1), "mrays_ao_closest": round(n * steps / (t[1] * 1e-3) / 1e6, 1),
            "mrays_ao_anyhit": round(n * steps / (t[2] * 1e-3) / 1e6, 1), "ms_per_step": round(sum(t) / steps, 4),
            "roofline": {"bound": "hbm", "kernel": "KTraceWide<closest> + exact-resolution tail", "achieved": round(closest_gbs, 1), "peak": peak,
                         "unit": "GB/s", "frac": round(closest_gbs / peak, 4), "bytes_per_ray": BYTES_CLOSEST}}


def e2e_plugin(sc, spp, rank, world, local, rays_per_step, barrier, max_over_ranks, args):
    """Config 3 through libTracerDLL_B200.so, driven through TracerI exactly as MRay's TracerThread would
    (oracle/ref_build/tracer_driver.cpp is that host application's stand-in; it contains none of the path's algorithms):
    scene upload from host arrays, CommitSurfaces (BVH build), StartRender, DoRenderWork until triggerSave with every
    film section copied to pinned host memory and accumulated by the caller. With N ranks every process renders its
    sample range (MRB_SPP_SHARD) and the host images are summed onto rank 0 inside the timed region."""
    import torch
    import torch.distributed as dist
    from mray_b200 import scenes
    plugin = os.path.join(ROOT, "mray_b200", "lib", "libTracerDLL_B200.so")
    O, bsc, alb = driver_scene(sc)
    if not (os.path.exists(plugin) and O.driver_available()):
        return {"value": None, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "unavailable": "libTracerDLL_B200.so / libtracer_driver.so were not prebuilt"}
    os.environ["MRB_DEVICE"] = str(local)
    os.environ["MRB_SPP_SHARD"] = "%d/%d" % (rank, world)
    s0, s1 = shard_samples(spp, world, rank)
    steps = max(2, min(args.steps, 3))

    def once():
        """One render through TracerI. Timed: everything the call does to produce the image on the host — scene upload
        calls, CommitSurfaces, StartRender, the DoRenderWork loop with every hand-off and the caller's accumulation, and
        (N > 1) the sum of the ranks' host images on rank 0. Not timed: tracer tear-down after the image exists."""
        img, w, st = O.driver_render(plugin, bsc, alb, len(sc["palb"]), sc["prad"], scenes.ARCADE_CAMERA, W, H, spp,
                                     renderer=RENDERER, sample_mode=SAMPLE_MODE, rr_range=RR, seed=0, burst_size=BURST)
        sec = st["scene_s"] + st["commit_s"] + st["start_s"] + st["render_s"]
        if world > 1:
            t0 = time.perf_counter()
            t = torch.from_numpy(np.concatenate([img * w[..., None], w[..., None]], axis=-1)).cuda()
            dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
            h = t.cpu().numpy()
            img, w = h[..., :3] / np.maximum(h[..., 3:], 1e-20), h[..., 3]
            sec += time.perf_counter() - t0
        return sec, st, w

    once()                                           # warm: library load, allocations, page-locking
    barrier()
    tt, st, w = 0.0, None, None
    for _ in range(steps):
        barrier()
        dt, st, w = once()
        tt += dt
    barrier()
    tt = max_over_ranks(tt)
    sec = tt / steps
    handoffs = int(st["iterations"])
    geometry = bsc["positions"].nbytes + bsc["normals"].nbytes + bsc["indices"].nbytes + (bsc["positions"].shape[0] * 8)
    return {"value": round(rays_per_step / sec / 1e6, 2), "unit": "Mrays/s",
            "ms_per_spp_1080p": round(sec * 1e3 / spp, 4), "ms_per_step": round(sec * 1e3, 1), "steps": steps,
            # per rank: geometry + attribute tables uploaded by the scene calls; film sections downloaded by the hand-offs
            "h2d_bytes_per_step": int(geometry + alb.nbytes + 4096), "d2h_bytes_per_step": int(handoffs * 4 * W * H * 4),
            "film_handoffs_per_step": handoffs, "commit_surfaces_ms": round(1e3 * st["commit_s"], 2),
            "start_render_ms": round(1e3 * st["start_s"], 2), "do_render_work_loop_ms": round(1e3 * st["render_s"], 1),
            "scene_upload_ms": round(1e3 * st.get("scene_s", 0.0), 1), "close_ms": round(1e3 * st.get("close_s", 0.0), 1),
            "driver_call_ms": round(1e3 * st.get("total_s", 0.0), 1),
            "film_weight_min_max": [float(w.min()), float(w.max())],
            "contract": "libTracerDLL_B200.so through TracerI: host scene arrays -> CommitSurfaces -> StartRender -> DoRenderWork "
                        "(renderMode Throughput, burstSize %d) until triggerSave; sections copied to pinned host memory and accumulated "
                        "in fp64 by the caller%s; rays = those of the device-timed leg (same seed, same paths)"
                        % (BURST, "; host images summed onto rank 0 over NCCL" if world > 1 else "")}


# ------------------------------------------------------------------------------------------------
# CPU side: the unmodified reference path tracer (oracle/_ref), all host threads, through TracerI
# ------------------------------------------------------------------------------------------------
def reference_dll():
    for name in ("libTracerDLL_CPU_rel.so", "libTracerDLL_CPU.so"):
        path = os.path.join(ROOT, "oracle", "_ref", name)
        if os.path.exists(path):
            return path, ("release flags" if "rel" in name else "-O2 IEEE flags")
    return None, None


def reference_render(O, bsc, alb, sc, width, height, spp):
    """One render of the reference's own (R)PathTracerSpectral on its CPU device backend; returns the driver's stats."""
    from mray_b200 import scenes
    dll, _ = reference_dll()
    _, wgt, st = O.driver_render(dll, bsc, alb, len(sc["palb"]), sc["prad"], scenes.ARCADE_CAMERA, width, height, spp,
                                 renderer=RENDERER, sample_mode=SAMPLE_MODE, rr_range=RR, seed=0, threads=0, host_exe=True,
                                 burst_size=1)
    return st


def ref_line(st, width, height, spp):
    paths = float(st["paths"])
    rays = paths * (REF_CLOSEST_PER_PATH + REF_NEE_PER_PATH)
    sec = float(st["render_s"])
    return rays / sec / 1e6, sec * 1e3 / spp * (W * H) / (width * height), paths / sec / 1e6


def reference_pt_sample(sc, budget_s=20.0):
    dll, flavour = reference_dll()
    if dll is None:
        return {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "reference", "sample": "oracle/_ref was never built"}
    O, bsc, alb = driver_scene(sc)
    w, h = W // 4, H // 4
    st = reference_render(O, bsc, alb, sc, w, h, 1)                          # calibration: 1/16 of the pixels, 1 spp
    per_px = float(st["render_s"]) / (w * h)
    scale = 2 if per_px * (W // 2) * (H // 2) <= budget_s else 4
    if scale == 2:
        w, h = W // 2, H // 2
        st = reference_render(O, bsc, alb, sc, w, h, 1)
    mrays, ms_spp, mpaths = ref_line(st, w, h, 1)
    return {"value": round(mrays, 4), "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "reference",
            "ms_per_spp_1080p": round(ms_spp, 1), "mpaths_s": round(mpaths, 4), "bvh_build_ms": round(1e3 * st["commit_s"], 1),
            "sample": "the unmodified reference (R)PathTracerSpectral, CPU device backend (%s), all host threads, through TracerI on "
                      "the same scene and settings at %dx%d, 1 spp (%.1f s); ms/spp scaled by the pixel ratio; rays = paths x "
                      "%.2f (closest %.2f + NEE shadow %.2f per path, measured on the same estimator)"
                      % (flavour, w, h, st["render_s"], REF_CLOSEST_PER_PATH + REF_NEE_PER_PATH, REF_CLOSEST_PER_PATH, REF_NEE_PER_PATH)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dll, flavour = reference_dll()
    if dll is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref was never built (needs /root/reference at build time)"}), flush=True)
        return
    sc = build_scene()
    O, bsc, alb = driver_scene(sc)
    w, h = W // 4, H // 4                                                     # bounded sample: 1/16 of the pixels, 1 spp per step
    times, paths = [], 0.0
    st = None
    for k in range(args.warmup + args.steps):
        st = reference_render(O, bsc, alb, sc, w, h, 1)
        if k >= args.warmup:
            times.append(float(st["render_s"])); paths += float(st["paths"])
    sec = sum(times)
    rays = paths * (REF_CLOSEST_PER_PATH + REF_NEE_PER_PATH)
    value = rays / sec / 1e6
    ms_spp = 1e3 * sec / len(times) * (W * H) / (w * h)
    sample = ("the unmodified reference (R)PathTracerSpectral, CPU device backend (%s), all host threads, through TracerI; each step = "
              "%dx%d at 1 spp of the same scene and settings (DoRenderWork loop wall time); ms/spp scaled by the pixel ratio; rays = paths "
              "x %.2f (closest %.2f + NEE shadow %.2f per path)"
              % (flavour, w, h, REF_CLOSEST_PER_PATH + REF_NEE_PER_PATH, REF_CLOSEST_PER_PATH, REF_NEE_PER_PATH))
    out = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": "Mrays/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * sec / len(times), 3),
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "ms_per_spp_1080p": round(ms_spp, 1), "mpaths_s": round(paths / sec / 1e6, 4),
                      "bvh_build_ms": round(1e3 * st["commit_s"], 1)},
           "cpu_baseline": {"value": round(value, 4), "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "reference", "sample": sample},
           "e2e": {"value": round(value, 4), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp", type=int, default=SPP, help="samples per pixel of one step (the metric's configuration is 1024)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
