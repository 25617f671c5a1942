#!/usr/bin/env python
"""bench.py — BASELINE.json metric "Mrays/sec and ms/spp at 1080p (device-timed)" on config 2:
procedural Sponza-scale mesh (~264 K triangles), 1920x1080 pinhole primary rays (closest hit) +
ambient-occlusion batch (closest hit AND any hit) = 3 x 2 073 600 rays per step, BVH build timed
separately. One process per GPU; ranks are independent (scene + BVH replicated, every rank traces a
full 1080p batch with its own AO seed) — weak scaling, no data-path collective.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0). `value` is device-timed (CUDA events on the launching stream) with all
inputs resident in HBM; `e2e` goes through the host-pointer C-ABI calls (pinned host buffers, H2D/D2H
inside the timed region); `--impl reference` times the reference's own CPU code (oracle/_ref, built
from /root/reference) on the host cores on a bounded sample of the same workload.
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
BYTES_CLOSEST = 764  # 64 B stream I/O + ceil(log2 N)=18 x 36 B descent + 52 B leaf
BYTES_ANY = 736
METRIC = "Mrays/sec at 1080p (primary closest-hit + AO closest-hit + AO any-hit, device-timed)"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_workload(seed_offset=0):
    """Scene + the two ray batches. AO rays are derived from the primary hits, which come from the
    device path (parity with the oracle is the tests' job, not the bench's)."""
    from mray_b200 import scenes
    p, i = scenes.arcade_mesh()
    rays = scenes.pinhole_rays(W, H, **scenes.ARCADE_CAMERA)
    return p, i, rays


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
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import mray_b200
    from mray_b200 import capi, scenes

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL writes its version / debug lines to stdout by default; stdout carries exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = mray_b200.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream)

    p, i, rays_np = build_workload()
    n = rays_np.shape[0]
    # ---- BVH build (device timed inside the library, inputs resident in HBM) ----
    dp, di = torch.from_numpy(p).cuda(), torch.from_numpy(i.view(np.int32)).cuda()
    build_ms = []
    acc = None
    for _ in range(3):
        if acc is not None:
            acc.close()
        acc = mray_b200.Accelerator(ctx, dp, di)
        build_ms.append(float(acc.info.buildMs))
    build_best = min(build_ms)

    def new_outputs():
        return (torch.full((n, 4), -1, dtype=torch.int32, device="cuda"),
                torch.zeros((n, 2), dtype=torch.float32, device="cuda"))

    # primary hits -> AO rays (host side helper, untimed set-up)
    d_primary = torch.from_numpy(rays_np).cuda()
    keys, hits = new_outputs()
    work = d_primary.clone()
    acc.cast_rays(keys, hits, work, None, capi.MRB_TRACE_WIDE)
    torch.cuda.synchronize()
    prim = keys.cpu().numpy().view(np.uint32)[:, 0]
    tprim = work.cpu().numpy()[:, 7]
    e = acc.export_lbvh()
    diam = float(np.linalg.norm(e["accel_aabb"][3:] - e["accel_aabb"][:3]))
    ao_np = scenes.ao_rays(rays_np, prim, tprim, p, i, 0.15 * diam, seed_offset=rank * n)
    d_ao = torch.from_numpy(ao_np).cuda()
    words = (n + 31) // 32

    work_p, work_a = d_primary.clone(), d_ao.clone()
    keys_a, hits_a = new_outputs()
    bits = torch.full((words,), -1, dtype=torch.int32, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def reset():
        work_p.copy_(d_primary); work_a.copy_(d_ao)
        keys.fill_(-1); keys_a.fill_(-1); bits.fill_(-1)

    def step(timed):
        reset()  # untimed: restores tMax / outputs; also evicts the previous step's lines from L2
        ev[0].record(stream)
        acc.cast_rays(keys, hits, work_p, None, capi.MRB_TRACE_WIDE)
        ev[1].record(stream)
        acc.cast_rays(keys_a, hits_a, work_a, None, capi.MRB_TRACE_WIDE)
        ev[2].record(stream)
        acc.cast_visibility_rays(bits, d_ao, None, capi.MRB_TRACE_WIDE)
        ev[3].record(stream)
        if timed:
            torch.cuda.synchronize()
            return [ev[k].elapsed_time(ev[k + 1]) for k in range(3)]
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step(False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    t_primary = t_ao = t_any = 0.0
    for _ in range(args.steps):
        a, b, c = step(True)
        t_primary += a; t_ao += b; t_any += c
    barrier()
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    fallback = ctx.last_fallback_count
    total_ms = t_primary + t_ao + t_any
    if world > 1:
        tt = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms_max = float(tt.item())
    else:
        total_ms_max = total_ms
    rays_per_step = 3 * n
    value = world * rays_per_step * args.steps / (total_ms_max * 1e-3) / 1e6

    # ---- e2e: host-pointer C-ABI calls, pinned host buffers, copies inside the timed region ----
    def pinned(shape, dtype):
        t = torch.empty(shape, dtype=dtype).pin_memory()
        return t, t.numpy()

    hp_t, hp = pinned((n, 8), torch.float32); ha_t, ha = pinned((n, 8), torch.float32)
    hk_t, hk = pinned((n, 4), torch.int32); hh_t, hh = pinned((n, 2), torch.float32)
    hb_t, hb = pinned((words,), torch.int32)
    hk_u, hb_u = hk.view(np.uint32), hb.view(np.uint32)

    def e2e_step(mode):
        hp[:] = rays_np; ha[:] = ao_np; hk_u[:] = 0xFFFFFFFF; hb_u[:] = 0xFFFFFFFF  # host-side set-up, untimed
        t0 = time.perf_counter()
        acc.cast_rays(hk_u, hh, hp, None, mode)           # H2D rays (+ keys/hits), trace, D2H
        acc.cast_rays(hk_u, hh, ha, None, mode)
        acc.cast_visibility_rays(hb_u, ha, None, mode)
        checksum = int(hk_u[:, 0].sum(dtype=np.uint64)) ^ int(hb_u.sum(dtype=np.uint64))  # result read on host
        return time.perf_counter() - t0, checksum

    def e2e_measure(mode):
        for _ in range(2):
            e2e_step(mode)
        barrier()
        t = 0.0
        for _ in range(e2e_steps):
            dt, _ = e2e_step(mode)
            t += dt
        barrier()
        if world > 1:
            tt = torch.tensor([t], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        return world * rays_per_step * e2e_steps / t / 1e6
    e2e_steps = max(3, min(args.steps, 10))
    # headline: the caller does not need its output buffers read (MRB_TRACE_FRESH_OUTPUTS: misses get INVALID keys),
    # so only the rays travel host -> device; the reference's "untouched on miss" contract uploads keys and hits too
    e2e_value = e2e_measure(capi.MRB_TRACE_WIDE | capi.MRB_TRACE_FRESH_OUTPUTS)
    e2e_preserving = e2e_measure(capi.MRB_TRACE_WIDE)
    ray_b, key_b, hit_b = n * 32, n * 16, n * 8
    h2d = 3 * ray_b
    d2h = 2 * (ray_b + key_b + hit_b) + words * 4

    # ---- config 3 flavour: the full wavefront path tracer (NEE+MIS, rrRange [3,8]) at 1080p on the same mesh ----
    pt = None
    if os.environ.get("MRB_BENCH_SKIP_PT") is None:
        pidx, pranges, pkeys, palb, prad, _ = scenes.arcade_materials(p, i)
        pacc = mray_b200.Accelerator(ctx, dp, torch.from_numpy(pidx.view(np.int32)).cuda(), prim_ranges=pranges, light_or_mat_keys=pkeys)
        pt_spp = 8

        def run_pt(spectrum, partition=False, sampler="Independent"):
            pr = mray_b200.Renderer(ctx, pacc, p.shape[0], pidx.shape[0], palb, prad, scenes.ARCADE_CAMERA, W, H, pt_spp,
                                    sample_mode="WithNEEAndMIS", rr_range=(3, 8), seed=rank, partition_rays=partition, spectrum=spectrum,
                                    sampler=sampler)
            pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            pr.iterate(2); torch.cuda.synchronize()      # warm
            l0 = ctx.launch_count
            pe0.record(stream)
            while True:
                pr.iterate(8)
                pst = pr.stats()
                if pst.finished:
                    break
            pe1.record(stream); torch.cuda.synchronize()
            pms = pe0.elapsed_time(pe1)
            res = {"ms_per_spp_1080p": round(pms / pt_spp, 3), "mrays_s": round((pst.closestRays + pst.shadowRays) / pms / 1e3, 1),
                   "mpaths_s": round(pst.pathsCompleted / pms / 1e3, 1), "iterations": int(pst.iterations),
                   "gpu_launches": int(ctx.launch_count - l0)}
            pr.close()
            return res
        pt = {"workload": "arcade mesh, 64 Lambert + 200 emissive tris, WithNEEAndMIS rr[3,8], %dx%d, %d spp" % (W, H, pt_spp),
              "PathTracerRGB": run_pt(None),
              "PathTracerRGB_material_key_sort": run_pt(None, True),   # RayPartitioner on: not needed by the fused shading kernel
              "PathTracerRGB_ZSobol": run_pt(None, False, "ZSobol")}
        from mray_b200 import spectral
        if spectral.available():
            spec = mray_b200.Spectrum(ctx, spectral.load(), "HyperbolicPBRT")
            pt["PathTracerSpectral"] = run_pt(spec)      # config 3: hero-wavelength spectral transport, ACES_CG LUT
            spec.close()
        pacc.close()

    peak, peak_src = measured_peak()
    closest_bytes = 2 * n * BYTES_CLOSEST * args.steps
    achieved = closest_bytes / ((t_primary + t_ao) * 1e-3) / 1e9

    out = {
        "metric": METRIC, "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": round(total_ms_max / args.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": "config 2: procedural arcade mesh %d tris, %dx%d primary closest-hit + AO closest-hit + AO any-hit, "
                        "1 step = %d rays per GPU" % (i.shape[0], W, H, rays_per_step),
            "l2": "inputs+outputs per step (~330 MB) exceed the 126 MB L2; ray/tMax reset copies between timed regions",
            "ms_per_spp_1080p": round(total_ms_max / args.steps, 4),
            "mrays_primary": round(n * args.steps / (t_primary * 1e-3) / 1e6, 1),
            "mrays_ao_closest": round(n * args.steps / (t_ao * 1e-3) / 1e6, 1),
            "mrays_ao_anyhit": round(n * args.steps / (t_any * 1e-3) / 1e6, 1),
            "bvh_build_ms": round(build_best, 4), "bvh_build_mtris_s": round(i.shape[0] / (build_best * 1e-3) / 1e6, 1),
            "wide_nodes": int(acc.info.wideNodeCount), "exact_fallback_rays_last_cast": fallback,
            "parallelism": "independent ranks, BVH replicated" if world > 1 else "single GPU",
            "path_tracer_1080p": pt,
        },
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 2), "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "contract": "MRB_TRACE_FRESH_OUTPUTS (rays uploaded; rays, keys, hits and visibility bits downloaded)",
                "preserving_caller_outputs": {"value": round(e2e_preserving, 2), "h2d_bytes_per_step": 3 * ray_b + 2 * (key_b + hit_b) + words * 4}},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "KTraceWide<closest>", "achieved": round(achieved, 1), "peak": peak,
                     "unit": "GB/s", "frac": round(achieved / peak, 4),
                     # dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of the primary and the AO closest-hit
                     # launch in the ncu --set full capture profiles/r1_ktracewide_f.md (142.1 MB and 132.1 MB)
                     "traffic": 137.1e6, "algorithmic_bytes_per_launch": n * BYTES_CLOSEST,
                     "bytes_per_ray": BYTES_CLOSEST, "peak_source": peak_src},
    }
    if rank == 0:
        if world == 1:
            out["cpu_baseline"] = cpu_baseline(p, i, rays_np, ao_np, budget_s=12.0)
        print(json.dumps(out), flush=True)
    acc.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# CPU side: the reference's own code (oracle/_ref) or, if that was never built, the oracle port
# ------------------------------------------------------------------------------------------------
def cpu_tracer():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    rel = os.path.join(ROOT, "oracle", "_ref", "libref_taps_rel.so")
    par = os.path.join(ROOT, "oracle", "_ref", "libref_taps.so")
    if os.path.exists(rel) or os.path.exists(par):
        import ctypes as C
        lib = C.CDLL(rel if os.path.exists(rel) else par)
        f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS"); u32 = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
        u8 = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
        lib.ref_lbvh_trace.argtypes = [f32, C.c_uint32, u32, C.c_uint32, u32, f32, C.c_uint32, f32, C.c_uint32, C.c_int,
                                       C.c_int, u32, f32, f32, u8]
        cores = int(lib.ref_thread_count())

        def trace(p, i, b, rays, mode):
            n = rays.shape[0]
            prim = np.zeros(n, np.uint32); t = np.zeros(n, np.float32); bary = np.zeros((n, 2), np.float32); back = np.zeros(n, np.uint8)
            lib.ref_lbvh_trace(p, p.shape[0], i, i.shape[0], b.nodes, b.boxes, b.nodes.shape[0], rays, n, mode, 0, prim, t, bary, back)
            return prim
        flavour = "release flags" if os.path.exists(rel) else "-O2 IEEE flags"
        return "reference", cores, trace, O, ("reference TraverseLBVHStack + Ray::IntersectsAABB/IntersectsTriangle compiled from "
                                              "/root/reference (%s), all host threads" % flavour)
    cores = 1

    def trace(p, i, b, rays, mode):
        return O.oracle_trace(p, i, b, rays, mode)[0]
    return "port", cores, trace, O, "oracle/mray_oracle.c scalar port, 1 thread"


def cpu_baseline(p, i, rays, ao, budget_s=12.0):
    kind, cores, trace, O, desc = cpu_tracer()
    b = O.oracle_build(p, i)
    n = rays.shape[0]

    def sample(stride):
        sel = np.arange(0, n, stride)
        r0, r1 = np.ascontiguousarray(rays[sel]), np.ascontiguousarray(ao[sel])
        t0 = time.perf_counter()
        trace(p, i, b, r0, 0); trace(p, i, b, r1, 0); trace(p, i, b, r1, 1)
        return 3 * sel.size, time.perf_counter() - t0
    cnt, dt = sample(256)                      # calibration
    rate = cnt / max(dt, 1e-6)
    stride = max(1, int(np.ceil(3 * n / max(rate * budget_s, 1.0))))
    cnt, dt = sample(stride)
    return {"value": round(cnt / dt / 1e6, 4), "unit": "Mrays/s", "cores": cores, "kind": kind,
            "sample": "every %d-th ray of the 3 batches (%d rays, %.1f s); %s" % (stride, cnt, dt, desc)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    p, i, rays = build_workload()
    kind, cores, trace, O, desc = cpu_tracer()
    b = O.oracle_build(p, i)
    n = rays.shape[0]
    from mray_b200 import scenes
    # AO rays need primary hits: computed once with the CPU tracer on the sample only
    budget = 60.0 / max(1, args.steps + args.warmup)
    sel = np.arange(0, n, 256)
    t0 = time.perf_counter(); trace(p, i, b, np.ascontiguousarray(rays[sel]), 0); rate = sel.size / (time.perf_counter() - t0)
    stride = max(1, int(np.ceil(3 * n / max(rate * budget, 1.0))))
    sel = np.arange(0, n, stride)
    r0 = np.ascontiguousarray(rays[sel])
    prim, t, _, _ = O.oracle_trace(p, i, b, r0, 0)
    diam = float(np.linalg.norm(b.accel_aabb[3:] - b.accel_aabb[:3]))
    r1 = scenes.ao_rays(r0, prim, t, p, i, 0.15 * diam)
    times = []
    for k in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        trace(p, i, b, r0, 0); trace(p, i, b, r1, 0); trace(p, i, b, r1, 1)
        if k >= args.warmup:
            times.append(time.perf_counter() - t0)
    cnt = 3 * sel.size
    value = cnt * len(times) / sum(times) / 1e6
    sample = "every %d-th ray of the 3 batches (%d rays/step); %s" % (stride, cnt, desc)
    # config 3 side by side: the UNMODIFIED reference path tracer (its CPU backend, through TracerI) on the same mesh,
    # materials and lights at 1080p, 1 spp (wall time of its DoRenderWork loop, all host threads)
    ref_pt = None
    if os.environ.get("MRB_BENCH_SKIP_PT") is None and kind == "reference" and O.driver_available():
        try:
            pidx, pranges, pkeys, palb, prad, tri_mat = scenes.arcade_materials(p, i)
            mat = np.where(tri_mat < 0, len(palb), tri_mat).astype(np.uint32)
            bsc = O.batched_scene(p, pidx, mat)
            dll = os.path.join(ROOT, "oracle", "_ref", "libTracerDLL_CPU.so")
            alb = np.concatenate([palb, np.zeros((1, 3), np.float32)])
            # bounded sample: 1 spp at half resolution per axis (a quarter of the 1080p paths), scaled by 4
            _, wgt, st = O.driver_render(dll, bsc, alb, len(palb), prad, scenes.ARCADE_CAMERA, W // 2, H // 2, 1, renderer="PathTracerRGB",
                                         sample_mode="WithNEEAndMIS", rr_range=(3, 8), seed=0, threads=0, host_exe=True)
            ref_pt = {"workload": "arcade mesh, 64 Lambert + 200 emissive tris, WithNEEAndMIS rr[3,8]; sample: %dx%d, 1 spp, time x 4" % (W // 2, H // 2),
                      "PathTracerRGB": {"ms_per_spp_1080p": round(4e3 * st["render_s"], 1), "mpaths_s": round(st["paths"] / st["render_s"] / 1e6, 3),
                                        "bvh_build_ms": round(1e3 * st["commit_s"], 1), "iterations": st["iterations"]},
                      "kind": "reference TracerDLL (CPU backend) through TracerI", "cores": cores}
        except Exception as e:   # a baseline leg must not take the headline line down
            ref_pt = {"error": str(e)[:200]}
    out = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": "Mrays/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * sum(times) / len(times), 3),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "config 2: procedural arcade mesh %d tris, %dx%d primary closest-hit + AO closest-hit + "
                                  "AO any-hit (bounded sample)" % (i.shape[0], W, H),
                      "path_tracer_1080p": ref_pt},
           "cpu_baseline": {"value": round(value, 4), "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": round(value, 4), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
