#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 path-tracing hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config C2]

A "step" is one full render of the workload (BASELINE.json configs[1], "C2": the One-Weekend
random scene, 1200x675, 500 spp, depth 50 = 405 M samples) on every rank.  With N > 1 the path
is sample-split: rank r renders global samples [r*spp, (r+1)*spp) of every pixel,
the FP32 accumulators are summed with one NCCL reduce and rank 0 resolves the N*spp image, so
per-GPU work is fixed ("weak" scaling) and `value` counts the samples of all ranks.

Metrics:  value = Msamples/s with the scene resident in HBM (device time, CUDA events, max over
ranks); e2e = the same through the host-buffer plugin call (scene upload + render + RGB8 readback
every step, wall clock).  `--impl reference` times the CPU restatement of the reference's own
algorithm (oracle/: random-axis BVH + recursive radiance, all host threads) on a bounded sample.
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

from raytracer_go_b200 import scenes  # noqa: E402

# Algorithmic FP32 lane-instructions per unit of work, from the reference's own arithmetic
# (SURVEY.md §8d / DESIGN.md): one AABB slab test, one sphere discriminant, completing a hit,
# one scatter, one camera ray.
F_BOX, F_SPH, F_HIT, F_SHADE, F_GEN = 24, 24, 60, 80, 40
# Algorithmic bytes per unit: 32-byte node, 16-byte sphere, 32-byte material record.
B_BOX, B_SPH, B_HIT = 32, 16, 32


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])), mx.append(float(r[2])), power.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def workload(name, width=None, spp=None):
    scene, opts = scenes.build_config(name, width, spp)
    return scene, opts


def cpu_reference_rate(scene, cam_opts, target_seconds, threads=0):
    """Times the oracle's restatement of the reference algorithm (BVH.Hit over a reference-style
    random-axis tree, recursive GetColor) on a bounded sample of the workload: the same image at a
    reduced spp chosen to take about `target_seconds`."""
    import ctypes as C
    from oracle import pyoracle as orc
    from raytracer_go_b200 import abi
    threads = threads or orc.hardware_threads()

    def run(spp):
        o = abi.rt_camera_options.from_buffer_copy(bytes(cam_opts))
        o.spp = spp
        cam = orc.camera_from_options(o)
        _, _, st = orc.render(scene, cam, scenes.RENDER_SEED, mode=orc.MODE_REF_BVH, order=orc.ORDER_RECURSIVE,
                              threads=threads)
        return st, cam
    st, cam = run(1)
    rate = st.samples / max(st.seconds, 1e-9)
    spp = int(max(1, min(cam_opts.spp, round(rate * target_seconds / (cam.width * cam.height)))))
    if spp > 1:
        st, cam = run(spp)
    return {"msamples_s": st.samples / st.seconds / 1e6, "mrays_s": st.rays / st.seconds / 1e6, "spp": spp,
            "seconds": st.seconds, "threads": threads, "width": cam.width, "height": cam.height,
            "samples": int(st.samples)}


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement of the reference, all host threads, bounded sample."""
    if rank != 0:
        return
    scene, cam_opts = workload(args.config, args.width, args.spp)
    per_step = max(1.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    rates, rays, last = [], [], None
    for i in range(args.warmup + args.steps):
        r = cpu_reference_rate(scene, cam_opts, per_step)
        if i >= args.warmup:
            rates.append(r["msamples_s"]), rays.append(r["mrays_s"])
        last = r
    v = float(np.mean(rates)) if rates else 0.0
    sample = (f"{last['width']}x{last['height']} at {last['spp']} spp of {cam_opts.spp} "
              f"({last['samples'] / 1e6:.1f} M samples/step), same scene and camera")
    line = {
        "impl": "reference", "metric": "Msamples/s", "value": v, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * last["seconds"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "mrays_s": float(np.mean(rays)) if rays else 0.0,
        "config": config_dict(args, cam_opts, scene, world),
        "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": last["threads"], "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "C++ restatement of the Go path (oracle/oracle.cpp): Go is not installed in this image, "
                "so the reference itself cannot be built; not the Go binary",
    }
    print(json.dumps(line), flush=True)


def config_dict(args, cam_opts, scene, world):
    cam_w = cam_opts.image_width
    return {"workload": f"{args.config}: {scene.name} scene, {scene.n_objects()} hittables, {cam_w} px wide, "
                        f"{cam_opts.spp} spp{' in total' if args.split == 'strong' else '/GPU'}, depth {cam_opts.max_depth}",
            "spheres": int(len(scene.spheres)), "quads": int(len(scene.quads)), "width": int(cam_w),
            "spp_per_gpu": int(cam_opts.spp) // max(1, world) if args.split == "strong" else int(cam_opts.spp),
            "spp_total": int(cam_opts.spp) * (1 if args.split == "strong" else world),
            "max_depth": int(cam_opts.max_depth),
            "parallelism": "single GPU" if world == 1 else
                           (f"tile-split x{world}: interleaved scanlines at all {int(cam_opts.spp) * world} spp, one gather"
                            if args.split == "tile" else f"sample-split x{world} ({args.split})"),
            "l2": "per-pass radiance buffer (up to 1 GiB) and survivor queue exceed L2; the scene is shared-memory resident by design",
            "seed": scenes.RENDER_SEED}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(scenes.CONFIGS))
    ap.add_argument("--width", type=int, default=None, help="override image width (testing)")
    ap.add_argument("--spp", type=int, default=None, help="override samples per pixel (testing)")
    ap.add_argument("--split", default="weak", choices=["weak", "strong", "tile"],
                    help="weak: every rank renders the config's spp (default, per-GPU work fixed); "
                         "strong: the config's spp is divided over the ranks (BASELINE config C5); "
                         "tile: interleaved scanlines at world*spp samples (weak scaling, no reduction)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling sub-record")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # every rank builds its own BVH on the host: share the box's cores instead of oversubscribing them
    os.environ.setdefault("RT_B200_BVH_THREADS", str(max(1, min(16, (os.cpu_count() or 1) // max(1, world)))))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return 0

    import torch
    import torch.distributed as dist
    from raytracer_go_b200 import api, lib

    L = lib.load()  # raises if librt_b200.so is missing: there is no fallback
    if L.rt_device_count() < 1:
        raise RuntimeError("bench.py: no sm_100 device visible and librt_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    scene_data, cam_opts = workload(args.config, args.width, args.spp)
    cam = api.camera_from_options(cam_opts)
    W, H = cam.width, cam.height
    n_pix = W * H
    from raytracer_go_b200 import sharding
    rows = None
    if args.split == "strong":
        sample_offset, spp, total_spp = sharding.sample_split_strong(rank, world, cam.spp)
    elif args.split == "tile":
        # interleaved scanlines, every rank renders all world*spp samples of its rows (weak scaling:
        # the per-GPU work is that of the 1-GPU run); no reduction, one gather of the rows
        sample_offset, spp, total_spp = 0, world * cam.spp, world * cam.spp
        rows = sharding.row_split(rank, world, H)
    else:
        sample_offset, spp, total_spp = sharding.sample_split_weak(rank, world, cam.spp)

    sc = api.Scene(scene_data, local_rank)
    stream = torch.cuda.current_stream()
    sc.set_stream(stream.cuda_stream)
    my_rows = rows[1] if rows else H
    samples_per_rank = my_rows * W * spp
    accum = torch.empty(my_rows * W * 3, dtype=torch.float32, device="cuda")

    def step():
        st = sc.render_accum_device(cam, accum.data_ptr(), scenes.RENDER_SEED, sample_offset=sample_offset,
                                    sample_count=spp, rows=rows)
        if rows:
            full = sharding.gather_rows(accum.view(my_rows, W, 3), H, dst=0)  # tile-split exchange: one NCCL gather
        else:
            full = sharding.reduce_accumulators(accum, dst=0)  # sample-split exchange: one NCCL reduce over NVLink
        rgb = None
        if rank == 0:
            rgb = api.resolve_device(full.data_ptr(), W, H, total_spp, local_rank, stream.cuda_stream)
        return st, rgb

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    rays = mk_ms = 0.0
    launches = mk_launches = 0
    work_bytes = survivors = 0
    for _ in range(args.steps):
        st, rgb = step()
        rays += st.rays
        work_bytes += st.work_bytes
        survivors += st.survivors
        mk_ms += st.ms_megakernel
        launches += st.kernel_launches + (1 if rank == 0 else 0)
        mk_launches += st.megakernel_launches
    e1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    tms = torch.tensor([dev_ms, wall_ms, rays], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = tms.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tms.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms, wall_ms, rays = float(mx[0]), float(mx[1]), float(sm[2])
    ms_per_step = dev_ms / args.steps
    total_samples = n_pix * total_spp  # all ranks together
    value = total_samples / (ms_per_step * 1e-3) / 1e6
    mrays = rays / args.steps / (ms_per_step * 1e-3) / 1e6

    # ---- end-to-end through the host-buffer plugin call (what Camera.Render does): scene arrays
    # H2D + BVH build + render + RGB8 D2H, every step; all ranks render their sample range ----
    e2e = None
    if not args.no_e2e:
        def e2e_step():
            with api.Scene(scene_data, local_rank) as s2:
                rgb8, acc, _ = s2.render(cam, scenes.RENDER_SEED, sample_offset=sample_offset, sample_count=spp,
                                         want_accum=world > 1 and not rows, rows=rows)
            return rgb8
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / n_e2e
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        h2d = scene_data.nbytes() + 256
        d2h = my_rows * W * 3 + (n_pix * 12 if world > 1 and not rows else 0)
        e2e = {"value": total_samples / float(t[0]) / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": float(t[0]) * 1e3,
               "what": "rt_scene_create (host BVH build + upload) + rt_render (RGB8 to host) + rt_scene_destroy per step"}

    # ---- strong scaling sub-record (BASELINE config C5's plan on this config): the SAME frame, its spp divided over
    # the ranks, one NCCL reduce; t1 = the whole frame on rank 0 alone, timed in this process, so the efficiency
    # t1 / (N * tN) compares like with like.  e2e = scene upload + render of the rank's sample range + reduce +
    # RGB8 to the host, per step, wall clock. ----
    strong = None
    if args.split == "weak" and not args.no_strong:
        so, sspp, stot = sharding.sample_split_strong(rank, world, cam.spp)

        def strong_step(scene_handle, offset, count, everyone=True):
            if count > 0:
                scene_handle.render_accum_device(cam, accum.data_ptr(), scenes.RENDER_SEED, sample_offset=offset, sample_count=count)
            else:
                accum.zero_()
            full = sharding.reduce_accumulators(accum, dst=0) if everyone else accum
            if rank == 0:
                return api.resolve_device(full.data_ptr(), W, H, stot, local_rank, stream.cuda_stream)

        def timed(fn, n):
            fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            a.record(stream)
            for _ in range(n):
                fn()
            b.record(stream)
            barrier()
            t = torch.tensor([a.elapsed_time(b) / n, (time.perf_counter() - t0) * 1e3 / n], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0]), float(t[1])
        n_strong = max(1, min(args.steps, 5))
        tn_ms, _ = timed(lambda: strong_step(sc, so, sspp), n_strong)
        if world > 1:   # the whole frame on rank 0 alone (the other ranks wait at the barrier)
            t1_ms, _ = timed(lambda: strong_step(sc, 0, cam.spp, everyone=False) if rank == 0 else None, n_strong)
        else:
            t1_ms = tn_ms

        def strong_e2e_step():
            with api.Scene(scene_data, local_rank) as s2:
                s2.set_stream(stream.cuda_stream)
                strong_step(s2, so, sspp)
        _, e2e_ms = timed(strong_e2e_step, max(1, min(args.steps, 3)))
        frame = n_pix * cam.spp
        strong = {"what": f"the {cam.spp}-spp frame divided over {world} GPU(s) (sample-split, one NCCL reduce)",
                  "value": frame / (tn_ms * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": tn_ms,
                  "ms_per_step_1gpu": t1_ms, "efficiency_vs_1gpu": t1_ms / (world * tn_ms),
                  "e2e": {"value": frame / (e2e_ms * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": e2e_ms,
                          "h2d_bytes_per_step": int(scene_data.nbytes() + 256), "d2h_bytes_per_step": int(n_pix * 3)},
                  "spp_per_gpu": [sharding.sample_split_strong(r, world, cam.spp)[1] for r in range(world)]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- algorithmic work per sample, counted by the instrumented kernels on the same workload at a
    # reduced spp (per-ray statistics do not depend on spp).  The roofline's numerator is the work of the plain
    # top-down traversal of the binary tree — round 1's definition, 1 755 lane-instr/sample on C2 — counted on a
    # scene handle built WITHOUT the leaf-start chains and rendered WITHOUT the per-pixel candidate lists; what today's
    # kernels execute (fewer box tests: leaf start skips the ancestors of the leaf a ray leaves, camera rays test their
    # pixel's candidate list instead of walking the tree) is reported next to it as `executed`. ----
    cnt_spp = max(1, min(spp, 16))  # (16: the candidate lists are built from 16 spp on, so the count sees what the run executes)

    def count_work(scene_handle):
        stc = scene_handle.render_accum_device(cam, accum.data_ptr(), scenes.RENDER_SEED, sample_offset=0,
                                               sample_count=cnt_spp, flags=1, rows=rows)
        n_box, n_sph, n_hit = stc.box_tests / stc.rays, stc.sphere_tests / stc.rays, stc.hits / stc.rays
        seg = stc.rays / stc.samples
        return {"instr_per_sample": F_GEN + seg * (n_box * F_BOX + n_sph * F_SPH + n_hit * (F_HIT + F_SHADE)),
                "segments_per_sample": seg, "box_tests_per_ray": n_box, "sphere_tests_per_ray": n_sph, "hit_fraction": n_hit,
                "bytes_per_ray": n_box * B_BOX + n_sph * B_SPH + n_hit * B_HIT}
    executed = count_work(sc)
    # the reference algorithm's work: every ray traverses the tree top-down — no leaf start, no per-pixel candidate lists
    saved = {k: os.environ.get(k) for k in ("RT_B200_LEAF_START", "RT_B200_PIXEL_LISTS")}
    os.environ["RT_B200_LEAF_START"] = os.environ["RT_B200_PIXEL_LISTS"] = "0"
    with api.Scene(scene_data, local_rank) as sc_ref:
        sc_ref.set_stream(stream.cuda_stream)
        algo = count_work(sc_ref)
    for k, v in saved.items():
        if v is None:
            del os.environ[k]
        else:
            os.environ[k] = v
    instr_per_sample, seg = algo["instr_per_sample"], algo["segments_per_sample"]
    bytes_per_ray = algo["bytes_per_ray"]
    peaks, peak_src = measured_peaks()
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    peak_instr = sm_count * 128 * peaks["sm_max_mhz"] * 1e6 / 1e12  # T lane-instr/s
    mk_ms_per_launch = mk_ms / max(1, mk_launches)
    samples_per_launch = samples_per_rank * args.steps / max(1, mk_launches)
    achieved = instr_per_sample * samples_per_launch / (mk_ms_per_launch * 1e-3) / 1e12
    smem_bw = bytes_per_ray * (rays / world / args.steps) / (mk_ms / args.steps * 1e-3) / 1e9
    work_bytes_per_launch = work_bytes / max(1, mk_launches)
    roofline = {
        "bound": "fp32_issue", "kernel": "primary_stage_kernel + render_kernel<SPLIT> (timed together, per pass)",
        "achieved": achieved, "peak": peak_instr,
        "unit": "T lane-instr/s", "frac": achieved / peak_instr,
        # HBM bytes per pass (both kernels + the ordered reduce), from this run's own counters (rt_stats.work_bytes):
        # a 16-byte radiance record written and read per path, a 48-byte queue entry written and read per path that
        # survives its first segment, the accumulator read + written per pass.  ncu's dram__bytes of the profiled
        # pass agrees within 1 % (profiles/r02*_kernels_ncu_full.txt).
        "traffic": work_bytes_per_launch,
        "traffic_detail": {"unit": "bytes of HBM traffic per pass, derived from the run's counters",
                           "survivors_per_sample": survivors / max(1.0, samples_per_rank * args.steps),
                           "framebuffer_bytes": n_pix * 12,
                           "note": "SURVEY §8d's algorithmic HBM bytes for this regime are the framebuffer alone; the rest is "
                                   "the design's per-path radiance record + survivor queue"},
        "peak_source": f"{sm_count} SMs x 128 lanes x sm_max_mhz {peaks['sm_max_mhz']:.0f} ({peak_src} MEASURED_PEAKS.json clock)",
        "instr_per_sample": instr_per_sample, "segments_per_sample": seg, "box_tests_per_ray": algo["box_tests_per_ray"],
        "sphere_tests_per_ray": algo["sphere_tests_per_ray"], "hit_fraction": algo["hit_fraction"], "counted_at_spp": cnt_spp,
        "numerator": "top-down traversal of the binary tree by every ray (round 1's definition); `executed` = what the kernels run "
                     "with leaf start and per-pixel candidate lists",
        "executed": {k: executed[k] for k in ("instr_per_sample", "box_tests_per_ray", "sphere_tests_per_ray")},
        "frac_executed": executed["instr_per_sample"] / instr_per_sample * achieved / peak_instr,
        "ms_per_launch": mk_ms_per_launch, "launches": mk_launches,
        "kernel_share_of_step": mk_ms / dev_ms if world == 1 else None,
        # BVH-fetch regime: algorithmic bytes/ray x rays/s.  Served from shared memory when the
        # scene is staged there (C1/C2/C3/C5), from L1/L2 otherwise (C4: ~86 MB, L2-resident), never from HBM
        "scene_bytes_per_ray": bytes_per_ray, "scene_fetch_gbs": smem_bw,
        "scene_in_shared_memory": bool(sc.bvh_info().in_shared_memory),
        "scene_fetch_vs_hbm_peak": smem_bw / peaks["hbm_gbs"],
        "hbm": {"note": "HBM carries the per-pass work buffers (radiance records, survivor queue) and the framebuffer; "
                        "the scene is staged in shared memory",
                "achieved": work_bytes / args.steps / (ms_per_step * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s"},
    }

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference_rate(scene_data, cam_opts, 12.0)
        cpu = {"value": r["msamples_s"], "unit": "Msamples/s", "cores": r["threads"], "kind": "port",
               "mrays_s": r["mrays_s"],
               "sample": f"{r['width']}x{r['height']} at {r['spp']} spp of {spp} ({r['samples'] / 1e6:.1f} M samples, "
                         f"{r['seconds']:.1f} s), same scene/camera; C++ restatement of the Go path, not the Go binary"}

    line = {
        "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if args.split == "strong" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "mrays_s": mrays,
        "config": config_dict(args, cam_opts, scene_data, world),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu, "wall_ms_per_step": wall_ms / args.steps,
        "strong": strong,
    }
    print(json.dumps(line), flush=True)
    sc.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
