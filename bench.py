#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json: Msamples/s of the unidirectional path tracer).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|c4|c5|smoke|shipped]

A "step" = one batch of `--spp-per-step` iterations (Render calls) of the workload image through the hot path.
Default workload: C2 = Cornell box 1024x1024, depth 8 (the configuration the metric is quoted on; K steps of
128 spp -> 8 steps are the full 1024-spp config).  N > 1 (launched by torchrun, one rank per GPU): the image's
32x32 screen tiles are interleaved over the ranks (strong scaling); after every step the float3
accumulation framebuffer is reduced to rank 0 with one NCCL reduce over NVLink, inside the timed region.

Keys: value = whole-job Msamples/s, device-timed, inputs resident in HBM; e2e = the same through the public
C-ABI call with HOST buffers (camera in, tonemapped image out) inside the timed region; roofline = algorithmic
bytes/sample x samples / kernel time vs the measured HBM peak (MEASURED_PEAKS.json); cpu_baseline = the
reference's own kernel bodies (oracle/_ref/libref_host_fast.so, kind "reference") or the CPU oracle port on the
host cores for a bounded sample.  --impl reference times that CPU reference arm alone."""
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

# algorithmic bytes per sample B = R*(N*32 + P*36) + H*80 + 24, with R, N, P, H as executed by the
# reference traversal (counted by the oracle's restatement of it, tests/test_bench_constants.py re-measures them)
ALGO = {
    "c1": dict(R=8.66, N=13.99, P=8.02),
    "c2": dict(R=10.17, N=14.02, P=8.06),
    "c3": dict(R=13.37, N=13.62, P=7.25),
    "c4": dict(R=10.26, N=233.7, P=66.9),
    "c5": dict(R=12.92, N=10.68, P=4.01),
    "smoke": dict(R=13.09, N=17.82, P=11.28),      # SURVEY 8(f).3 scenes (heterogeneous media)
    "shipped": dict(R=15.80, N=11.75, P=7.70),
}


# ncu-measured DRAM traffic per sample, whole wavefront (see profiles/); filled in from the capture of the round
DRAM_BYTES_PER_SAMPLE = {"c2": 1684.0}     # profiles/r01s_launches_summary.txt: 21.2 GB over 12.6 Msamples
# warp instructions per pool slot and wavefront step (ncu smsp__inst_executed.sum of one k_trace_small + one k_shade launch
# over a 2^20-slot pool, profiles/r01z_ncu_wavefront_kernels.txt; r01w: 134.17e6 + 43.86e6): the kernels of C2 are
# instruction-issue bound (the scene lives in shared memory), so the JSON also carries the fraction of the SMs' issue
# rate they reach
WARP_INST_PER_SLOT_STEP = {"c2": (134.10e6 + 43.13e6) / float(1 << 20)}


def algo_bytes_per_sample(w):
    a = ALGO[w]
    H = a["R"] * 2.0 / 3.0
    return a["R"] * (a["N"] * 32 + a["P"] * 36) + H * 80 + 24


def make_scene(pt, name):
    if name == "c2":
        return pt.scenes.cornell_pt(1024, 1024, 8), "cornell_box 1024x1024 depth=8 lambertian+area-light (BASELINE configs[1])"
    if name == "c1":
        return pt.scenes.cornell_pt(256, 256, 4), "cornell_box 256x256 depth=4 (BASELINE configs[0])"
    if name == "c3":
        return pt.scenes.veach_standin(768, 576, 17), "veach_bidir materials/lights/camera over stand-in geometry 768x576 depth=17 (configs[2])"
    if name == "c4":
        return pt.scenes.random_triangles(1_000_000, 2048, 2048, 8), "1M random triangles + analytic HDRI 2048x2048 depth=8 (configs[3])"
    if name == "c5":
        return pt.scenes.cornell_vol_caustic(512, 512, 17), "cornell_box vol_caustic homogeneous medium vpt 512x512 depth=17 (configs[4])"
    if name == "smoke":
        return pt.scenes.cornell_smoke(1024, 1024, 8, 1), "cornell_box + heterogeneous smoke (ratio tracking) vpt 1024x1024 depth=8 (SURVEY 8(f).3)"
    if name == "shipped":
        return pt.scenes.cornell_shipped_smoke(1024, 1024, 17), "the reference's shipped cornell_box/scene.json (heterogeneous medium) vpt 1024x1024 depth=17"
    raise SystemExit(f"unknown workload {name}")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows = []
        self.stop = False
        self.index = index
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        self.p = p
        for line in p.stdout:
            if self.stop:
                break
            self.rows.append([x.strip() for x in line.split(",")])
        p.kill()

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        try:
            self.p.kill()
        except Exception:
            pass

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = sorted(sm)[len(sm) // 2:]          # upper half = samples under load
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference(scene, seconds_target, threads=0, fast=True):
    """The reference's own Path/Volpath kernel bodies on the host cores (oracle/_ref), else the oracle port."""
    from tests import refhost
    cores = threads or (os.cpu_count() or 1)
    if refhost.have("libref_host_fast.so"):
        ref, kind = refhost.RefHost(fast=fast), "reference"
        run = lambda first, spp: ref.render(scene, first, spp, threads=cores)
    else:
        from tests.oracle_lib import Oracle
        orc, kind = Oracle(), "port"
        run = lambda first, spp: orc.render(scene, first, spp, threads=cores)
    t0 = time.time(); run(1, 1); t1 = time.time() - t0
    spp = max(1, min(64, int(seconds_target / max(t1, 1e-3))))
    t0 = time.time(); run(2, spp); dt = time.time() - t0
    n = scene.width * scene.height * spp
    return {"value": n / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": kind,
            "sample": f"{spp} iterations of {scene.width}x{scene.height} ({n / 1e6:.1f} Msamples) in {dt:.1f}s"}, n / dt / 1e6, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--spp-per-step", type=int, default=128)
    ap.add_argument("--pool", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))

    import gpu_pathtracer_b200 as pt
    scene, desc = make_scene(pt, a.workload)
    W, H = scene.width, scene.height
    config = {"workload": desc, "width": W, "height": H, "max_depth": scene.max_depth, "spp_per_step": a.spp_per_step,
              "prims": int(len(scene.prims)), "bvh_nodes": int(len(scene.nodes)),
              "l2_policy": "path pool + sample planes streamed per step exceed L2 (126 MB); no explicit flush",
              "parallelism": f"tile-sharded x{world}" if world > 1 else "single GPU"}

    if a.impl == "reference":
        if rank != 0:
            return
        # reference arm: the reference's own CPU implementation of the path on the host cores, bounded sample
        per_step = 20.0 / max(1, a.steps + a.warmup)
        vals = []
        for i in range(a.warmup + a.steps):
            cb, v, dt = cpu_reference(scene, per_step)
            if i >= a.warmup:
                vals.append((v, dt))
        v = float(np.mean([x[0] for x in vals])); ms = float(np.mean([x[1] for x in vals]) * 1e3)
        cb["value"] = v
        print(json.dumps({"impl": "reference", "metric": "Msamples/s", "value": v, "unit": "Msamples/s", "n_gpus": 0, "steps": a.steps,
                          "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic", "config": config, "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    shard = (rank, world, 32, 32) if world > 1 else None
    r = pt.PathTracer(scene, device=local, shard=shard, pool=a.pool or None)
    spp = a.spp_per_step
    npix = W * H
    acc_t = torch.empty(0)
    if world > 1:
        # wrap the library's accumulation framebuffer (device memory owned by the context) as a torch tensor
        import ctypes
        ptr = r.accum_device_ptr()
        class _Holder:  # noqa: E306
            __cuda_array_interface__ = {"shape": (npix * 3,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        acc_t = torch.as_tensor(_Holder(), device=f"cuda:{local}")
    out_dev = torch.empty(npix * 3, dtype=torch.float32, device=f"cuda:{local}")
    out_host_t = torch.empty((H, W, 3), dtype=torch.float32, pin_memory=True)              # pinned host image for the e2e arm
    out_host = out_host_t.numpy()
    full_acc = torch.empty(npix * 3, dtype=torch.float32, device=f"cuda:{local}") if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_and_tonemap(last_iter):
        """N > 1: ONE NCCL reduce of the float3 accumulation framebuffer per spp batch (tiles are disjoint, so the sum has
        one non-zero contributor per pixel), then Output's tonemap of the full image on rank 0."""
        full_acc.copy_(acc_t)                     # keep the context's own buffer shard-only for the next batch
        dist.reduce(full_acc, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            torch.cuda.current_stream().synchronize()       # the tonemap runs on the library's stream
            r.tonemap_device(full_acc.data_ptr(), last_iter, out_dev.data_ptr())

    def step_device(i):
        """inputs resident: camera struct is the only host->device traffic (104 B, like the reference's Render)."""
        r.render(1 + i * spp, reset=(i == 0), spp=spp, output=out_dev.data_ptr(), output_is_device=True)
        if world > 1:
            reduce_and_tonemap((i + 1) * spp)
        return r.stats()

    def step_e2e(i):
        """public call with HOST buffers: camera from host, tonemapped float3 image back to (pinned) host memory every step."""
        if world == 1:
            return r.render(1 + i * spp, reset=(i == 0), spp=spp, output=out_host)
        r.render(1 + i * spp, reset=(i == 0), spp=spp, output=out_dev.data_ptr(), output_is_device=True)
        reduce_and_tonemap((i + 1) * spp)
        if rank == 0:
            out_host_t.view(-1).copy_(out_dev)
            torch.cuda.synchronize()
        return out_host

    # ---- device-timed arm
    for i in range(a.warmup):
        step_device(i)
    barrier()
    launches = rays = dev_ms = 0.0
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        for i in range(a.steps):
            st = step_device(i)
            launches += st["launches"]; rays += st["rays"]; dev_ms += st["device_ms"]
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = clk.summary()
    # max over ranks of the device time (CUDA events on the context's stream) and of the wall time
    t = torch.tensor([dev_ms, wall_ms, launches, rays], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, wall_ms = float(tmax[0]), float(tmax[1]); launches, rays = float(tsum[2]), float(tsum[3])
    samples = float(npix) * spp * a.steps
    ms_per_step = wall_ms / a.steps
    value = samples / wall_ms / 1e3

    # ---- end-to-end arm (host buffers inside the timed region)
    step_e2e(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(a.steps):
        step_e2e(i)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t[0])
    e2e_val = samples / e2e_ms / 1e3

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        bps = algo_bytes_per_sample(a.workload)
        kernel_s = dev_ms / 1e3                       # wavefront kernels (trace+shade+resolve) of all timed steps, CUDA events
        achieved = bps * samples / world / kernel_s / 1e9 if kernel_s > 0 else 0.0
        # measured DRAM bytes per sample (ncu dram__bytes_read+write summed over every kernel of one bench step,
        # profiles/<round>_dram_bytes_*.txt) x samples of one step, per GPU; None for workloads without a capture
        dram_bps = DRAM_BYTES_PER_SAMPLE.get(a.workload)
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": (dram_bps * float(npix) * spp / world) if dram_bps else None,
                    "peak_source": "measured" if peaks else "fallback",
                    "algorithmic_bytes_per_sample": bps, "algorithmic_bytes_per_step": bps * float(npix) * spp / world,
                    "measured_dram_bytes_per_sample": dram_bps,
                    "kernel": "wavefront step = k_shade + k_trace(_small) + k_resolve of one spp batch (per GPU); the per-sample "
                              "figure of SURVEY 8(d) spans all three, so they are timed together with CUDA events"}
        if a.workload in WARP_INST_PER_SLOT_STEP and clocks.get("sm_mhz"):
            # slot-steps = (shade, trace) launch pairs x slots per lane; launches counts both kernels of every lane
            slots_per_lane = r.info("pool_per_lane")
            slot_steps = launches / 2.0 / world * slots_per_lane
            inst = WARP_INST_PER_SLOT_STEP[a.workload] * slot_steps
            peak_issue = 148 * 4 * clocks["sm_mhz"] * 1e6
            roofline["issue"] = {"warp_inst_per_sample": inst / (samples / world), "achieved_warp_inst_per_s": inst / kernel_s,
                                 "peak_warp_inst_per_s": peak_issue, "frac": inst / kernel_s / peak_issue,
                                 "note": "148 SMs x 4 schedulers x SM clock; instruction counts from ncu (profiles/r01z_ncu_wavefront_kernels.txt)"}
        cb = None
        if not a.no_cpu_baseline and world == 1:
            cb, _, _ = cpu_reference(scene, 12.0)          # bounded sample of the SAME workload on the host cores
            cb["sample"] += " of " + desc
        line = {"metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "device_ms_per_step": dev_ms / a.steps,
                "e2e": {"value": e2e_val, "unit": "Msamples/s", "h2d_bytes_per_step": 104, "d2h_bytes_per_step": npix * 12},
                "gpu_launches": int(launches), "rays_per_sample": rays / samples, "clocks": clocks, "roofline": roofline, "cpu_baseline": cb}
        print(json.dumps(line))
    r.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
