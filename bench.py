#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json: Msamples/s of the unidirectional path tracer).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|c4|c5|smoke|shipped]

A "step" = one batch of `spp_per_step` iterations (Render calls) of the workload image through the hot path.
Default workload: C2 = Cornell box 1024x1024, depth 8 (the configuration the metric is quoted on); spp_per_step is 128 per
GPU — 128 at N = 1 (8 steps are the full 1024-spp config), 128 x N at N GPUs, so the timed region does not shrink as
GPUs are added (reported as "scaling": "weak": per-GPU work per step is fixed, the image and the scene are the same).
N > 1 (launched by torchrun, one rank per GPU): the image's 16x16 screen tiles are dealt over the ranks;
after every step the float3 accumulation framebuffers are summed onto rank 0 by ONE NCCL reduce over NVLink — inside the
library (b200pt_render_reduce; torch.distributed only ships the 128-byte NCCL id and the timing scalars) and inside the
timed region.

Keys: value = whole-job Msamples/s, device-resident arm (device output pointer); e2e = the same through the public C-ABI
call with HOST buffers (camera in, tonemapped image out) inside the timed region; roofline = what bounds the dominant
kernel, from per-sample hardware counters measured with ncu (profiles/r02_counters.json, scripts/ncu_counters.py) x the
samples of the timed region / CUDA-event time: for scenes that live in shared memory the binding resource is
instruction issue (frac = useful-lane issue fraction; the HBM figures are reported next to it), for C4 it is
HBM / L2; cpu_baseline = the reference's own kernel bodies (oracle/_ref/libref_host_fast.so) on the host cores for a
bounded sample; reference_cuda = the reference's own CUDA integrator (oracle/_ref/libref_cuda.so) on
the same GPU, timed in a subprocess outside the timed region; extra.shipped_scene = a short leg on the reference's own
scene.json (heterogeneous medium); extra.c4 = a short leg on BASELINE configs[3]
(1 M triangles, 2048x2048).  --impl reference times the CPU reference arm alone."""
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
COUNTERS = os.path.join(ROOT, "profiles", "r02_counters.json")

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
    "shipped512": dict(R=15.80, N=11.75, P=7.70),
}


def algo_bytes_per_sample(w):
    a = ALGO[w]
    H = a["R"] * 2.0 / 3.0
    return a["R"] * (a["N"] * 32 + a["P"] * 36) + H * 80 + 24


def make_scene(pt, name, prep=None):
    if name == "c2":
        return pt.scenes.cornell_pt(1024, 1024, 8, prep=prep), "cornell_box 1024x1024 depth=8 lambertian+area-light (BASELINE configs[1])"
    if name == "c1":
        return pt.scenes.cornell_pt(256, 256, 4, prep=prep), "cornell_box 256x256 depth=4 (BASELINE configs[0])"
    if name == "c3":
        return pt.scenes.veach_standin(768, 576, 17, prep=prep), "veach_bidir materials/lights/camera over stand-in geometry 768x576 depth=17 (configs[2])"
    if name == "c4":
        return pt.scenes.random_triangles(1_000_000, 2048, 2048, 8, prep=prep), "1M random triangles + analytic HDRI 2048x2048 depth=8 (configs[3])"
    if name == "c5":
        return pt.scenes.cornell_vol_caustic(512, 512, 17, prep=prep), "cornell_box vol_caustic homogeneous medium vpt 512x512 depth=17 (configs[4])"
    if name == "smoke":
        return pt.scenes.cornell_smoke(1024, 1024, 8, 1, prep=prep), "cornell_box + heterogeneous smoke (ratio tracking) vpt 1024x1024 depth=8 (SURVEY 8(f).3)"
    if name == "shipped":
        return pt.scenes.cornell_shipped_smoke(1024, 1024, 17, prep=prep), "the reference's shipped cornell_box/scene.json (heterogeneous medium) vpt 1024x1024 depth=17"
    if name == "shipped512":
        return pt.scenes.cornell_shipped_smoke(512, 512, 17, prep=prep), "the reference's shipped cornell_box/scene.json (heterogeneous medium) vpt 512x512 depth=17"
    raise SystemExit(f"unknown workload {name}")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows = []
        self.stop = False
        self.index = index
        self.p = None
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


def reference_cuda_subprocess(workload, spp, timeout=600):
    """the reference's OWN CUDA integrator (oracle/_ref/libref_cuda.so: its pathtracer.cu compiled for
    sm_100a) on the same GPU and the same scene arrays — Render(iter) x spp with CUDA events around the loop — in a
    subprocess of its own, outside every timed region of this benchmark."""
    code = (
        "import sys, json; sys.path.insert(0, %r)\n"
        "import bench, gpu_pathtracer_b200 as pt\n"
        "from tests import refhost\n"
        "s, _ = bench.make_scene(pt, %r)\n"
        "ref = refhost.RefCuda(); ref.begin(s); ref.render(1, 2, want_output=False)\n"
        "_, ms = ref.render(1, %d, want_output=False); ref.end()\n"
        "n = s.width * s.height * %d\n"
        "print('REFCUDA ' + json.dumps({'value': n / ms / 1e3, 'unit': 'Msamples/s', 'ms': ms, 'spp': %d, 'samples': n}))\n"
    ) % (ROOT, workload, spp, spp, spp)
    try:
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=timeout, cwd=ROOT).stdout
        for ln in out.splitlines():
            if ln.startswith("REFCUDA "):
                return json.loads(ln[8:])
    except Exception as e:                                                   # noqa: BLE001
        return {"unavailable": str(e)[:200]}
    return {"unavailable": "no result line"}


def load_counters(workload):
    try:
        return json.load(open(COUNTERS)).get(workload)
    except Exception:
        return None


def roofline_record(workload, counters, samples_per_gpu, kernel_s, clocks, peaks):
    """What bounds the dominant kernel.  Per-sample counts come from the committed ncu capture; time is measured live."""
    peak_hbm = float(peaks.get("hbm_gbs", 6650.0))
    bps = algo_bytes_per_sample(workload)
    algo_gbs = bps * samples_per_gpu / kernel_s / 1e9 if kernel_s > 0 else 0.0
    rec = {"peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
           "algorithmic_bytes_per_sample": bps, "algorithmic_gbs": algo_gbs,
           "counters_file": os.path.relpath(COUNTERS, ROOT) if counters else None}
    if not counters:
        rec.update({"bound": "hbm", "achieved": algo_gbs, "peak": peak_hbm, "unit": "GB/s", "frac": min(algo_gbs / peak_hbm, 1.0), "traffic": None,
                    "note": "no ncu counters for this workload: algorithmic bytes only (clamped: the scene may be served from on-chip memory)"})
        return rec
    dram_bps = counters["dram_bytes_per_sample"]
    dram_gbs = dram_bps * samples_per_gpu / kernel_s / 1e9
    hbm = {"measured_dram_bytes_per_sample": dram_bps, "measured_dram_gbs": dram_gbs, "frac_of_peak": dram_gbs / peak_hbm, "peak": peak_hbm}
    sm_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0
    peak_issue = 148 * 4 * sm_mhz * 1e6
    wi = counters["warp_inst_per_sample"] * samples_per_gpu / kernel_s
    lanes = counters["active_lanes_per_inst"]
    issue = {"warp_inst_per_sample": counters["warp_inst_per_sample"], "active_lanes_per_inst": lanes, "achieved_warp_inst_per_s": wi,
             "peak_warp_inst_per_s": peak_issue, "issue_frac": wi / peak_issue, "useful_lane_frac": wi / peak_issue * lanes / 32.0,
             "note": "148 SMs x 4 schedulers x SM clock under load"}
    rec["dominant_kernel"] = counters.get("dominant_kernel")
    rec["kernel_time_shares"] = {k: round(v["time_share"], 4) for k, v in counters.get("kernels", {}).items()}
    if hbm["frac_of_peak"] >= issue["issue_frac"]:
        rec.update({"bound": "hbm", "achieved": dram_gbs, "peak": peak_hbm, "unit": "GB/s", "frac": dram_gbs / peak_hbm})
    else:
        # the scene is served from shared memory / L2: instruction issue binds; frac counts only the lanes that do work
        rec.update({"bound": "issue", "achieved": wi * lanes / 32.0 / 1e9, "peak": peak_issue / 1e9, "unit": "G useful warp-inst/s",
                    "frac": issue["useful_lane_frac"]})
    rec["traffic"] = dram_bps * samples_per_gpu          # measured DRAM bytes of the timed kernels (per GPU)
    rec["hbm"] = hbm
    rec["issue"] = issue
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--spp-per-step", type=int, default=0, help="iterations per step and GPU (default 128; c4: 8)")
    ap.add_argument("--pool", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the reference_cuda sub-record and the C4 leg")
    a = ap.parse_args()
    # only the JSON line goes to the real stdout: everything else this process prints at C level (the reference's
    # Scene::Init / BeginRender chatter in the baseline legs) is routed to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))

    import gpu_pathtracer_b200 as pt

    if a.impl == "reference":
        if rank != 0:
            return
        # reference arm: the reference's own CPU implementation of the path on the host cores, bounded sample; the scene is
        # prepared by the reference's own Scene::Init / Camera constructor too (tests.refhost.RefPrep) — nothing of
        # libb200pt.so is loaded in this arm
        from tests import refhost
        prep = refhost.RefPrep() if refhost.have("libref_host.so") else None
        scene, desc = make_scene(pt, a.workload, prep=prep)
        spp = a.spp_per_step or (8 if a.workload == "c4" else 128)
        config = {"workload": desc, "width": scene.width, "height": scene.height, "max_depth": scene.max_depth, "spp_per_step": spp,
                  "prims": int(len(scene.prims)), "bvh_nodes": int(len(scene.nodes)), "parallelism": "host cores (OpenMP over pixels)"}
        per_step = 20.0 / max(1, a.steps + a.warmup)
        vals = []
        for i in range(a.warmup + a.steps):
            cb, v, dt = cpu_reference(scene, per_step)
            if i >= a.warmup:
                vals.append((v, dt))
        v = float(np.mean([x[0] for x in vals])); ms = float(np.mean([x[1] for x in vals]) * 1e3)
        cb["value"] = v
        from gpu_pathtracer_b200 import _lib
        emit({"impl": "reference", "metric": "Msamples/s", "value": v, "unit": "Msamples/s", "n_gpus": 0, "steps": a.steps,
                          "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic", "config": config, "cpu_baseline": cb,
                          "product_library_loaded": _lib._lib is not None,
                          "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_workload(workload, steps, warmup, spp_gpu, want_e2e=True):
        """One workload through both arms; returns the fields of a bench line (rank 0) — device-timed `value`, `e2e`, counters."""
        scene, desc = make_scene(pt, workload)
        W, H = scene.width, scene.height
        npix = W * H
        spp = spp_gpu * world                                  # iterations per step: fixed per-GPU work
        shard = (rank, world, 16, 16) if world > 1 else None      # small tiles: every rank sees a fair sample of the image
        r = pt.PathTracer(scene, device=local, shard=shard, pool=a.pool or None)
        if world > 1:
            # the ONE collective lives in the library: ship the NCCL id (128 bytes) with torch.distributed, then
            # b200pt_comm_init on every rank
            ids = [pt.PathTracer.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            r.comm_init(world, rank, ids[0])
        out_dev = torch.empty(npix * 3, dtype=torch.float32, device=f"cuda:{local}")
        out_host_t = torch.empty((H, W, 3), dtype=torch.float32, pin_memory=True)          # pinned host image for the e2e arm
        out_host = out_host_t.numpy()

        def step_device(i):
            """inputs resident: the camera struct is the only host->device traffic (104 B, like the reference's Render)."""
            if world == 1:
                r.render(1 + i * spp, reset=(i == 0), spp=spp, output=out_dev.data_ptr(), output_is_device=True)
            else:
                r.render_reduce(1 + i * spp, reset=(i == 0), spp=spp, root=0, output=out_dev.data_ptr() if rank == 0 else None,
                                output_is_device=True)
            return r.stats()

        def step_e2e(i):
            """public call with HOST buffers: camera from host, tonemapped float3 image back to (pinned) host memory every step."""
            if world == 1:
                return r.render(1 + i * spp, reset=(i == 0), spp=spp, output=out_host)
            return r.render_reduce(1 + i * spp, reset=(i == 0), spp=spp, root=0, output=out_host if rank == 0 else None)

        for i in range(warmup):
            step_device(i)
        barrier()
        launches = rays = dev_ms = 0.0
        with ClockSampler(local) as clk:
            t0 = time.perf_counter()
            for i in range(steps):
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
        samples = float(npix) * spp * steps
        res = {"desc": desc, "scene": scene, "samples": samples, "spp": spp, "value": samples / wall_ms / 1e3, "ms_per_step": wall_ms / steps,
               "dev_ms": dev_ms, "launches": launches, "rays": rays, "clocks": clocks, "npix": npix, "fused": r.info("fused"), "lanes": r.info("lanes")}
        if want_e2e:
            step_e2e(0)
            barrier()
            t0 = time.perf_counter()
            for i in range(steps):
                step_e2e(i)
            barrier()
            e2e_ms = (time.perf_counter() - t0) * 1e3
            t = torch.tensor([e2e_ms], dtype=torch.float64, device=f"cuda:{local}")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res["e2e"] = samples / float(t[0]) / 1e3
        r.close()
        return res

    spp_gpu = a.spp_per_step or (8 if a.workload == "c4" else 128)
    res = run_workload(a.workload, a.steps, a.warmup, spp_gpu)
    scene = res["scene"]
    extra = {}
    if not a.no_extra and a.workload == "c2":
        # BASELINE configs[3], the HBM-meaningful configuration, rides along as a short leg at every N
        c4 = run_workload("c4", 3, 3, 8)
        if rank == 0:
            k_s = c4["dev_ms"] / 1e3
            extra["c4"] = {"workload": c4["desc"], "value": c4["value"], "e2e": c4.get("e2e"), "unit": "Msamples/s", "steps": 3, "warmup": 3,
                           "spp_per_step": c4["spp"], "ms_per_step": c4["ms_per_step"], "rays_per_sample": c4["rays"] / c4["samples"],
                           "roofline": roofline_record("c4", load_counters("c4"), c4["samples"] / world, k_s, c4["clocks"], peaks)}

        # ... and the scene the reference ships (cornell_box/scene.json: heterogeneous smoke, depth 17), SURVEY 8(f).3
        het = run_workload("shipped512", 3, 3, 16, want_e2e=False)
        if rank == 0:
            extra["shipped_scene"] = {"workload": het["desc"], "value": het["value"], "unit": "Msamples/s", "steps": 3, "warmup": 3,
                                      "spp_per_step": het["spp"], "ms_per_step": het["ms_per_step"], "rays_per_sample": het["rays"] / het["samples"],
                                      "roofline": roofline_record("shipped512", load_counters("shipped512"), het["samples"] / world, het["dev_ms"] / 1e3, het["clocks"], peaks)}
            if world == 1:
                rc = reference_cuda_subprocess("shipped512", 16)
                if "value" in rc:
                    rc["ratio_device"] = het["value"] / rc["value"]
                extra["shipped_scene"]["reference_cuda"] = rc

    if rank == 0:
        kernel_s = res["dev_ms"] / 1e3                   # wavefront kernels of all timed steps, CUDA events on the context's stream
        roofline = roofline_record(a.workload, load_counters(a.workload), res["samples"] / world, kernel_s, res["clocks"], peaks)
        roofline["kernel"] = ("k_wave_small: CTA-local wavefront, one persistent launch per spp batch (+ k_resolve)" if res["fused"] else
                              "wavefront step = k_shade + k_trace of one spp batch (+ k_resolve), timed together with CUDA events")
        cb = None
        ref_cuda = None
        if world == 1 and not a.no_cpu_baseline:
            cb, _, _ = cpu_reference(scene, 12.0)          # bounded sample of the SAME workload on the host cores
            cb["sample"] += " of " + res["desc"]
        if not a.no_extra and a.workload in ("c1", "c2", "c3", "c5"):
            ref_cuda = reference_cuda_subprocess(a.workload, 64 if a.workload == "c2" else 32)
            if "value" in ref_cuda:
                ref_cuda["ratio_device"] = res["value"] / ref_cuda["value"]
                ref_cuda["ratio_e2e"] = res["e2e"] / ref_cuda["value"]
                ref_cuda["what"] = "the reference's pathtracer.cu (Path/Volpath megakernel) compiled for sm_100a, same GPU, same scene arrays, 1 GPU"
        config = {"workload": res["desc"], "width": scene.width, "height": scene.height, "max_depth": scene.max_depth,
                  "spp_per_step": res["spp"], "spp_per_step_per_gpu": spp_gpu, "prims": int(len(scene.prims)), "bvh_nodes": int(len(scene.nodes)),
                  "l2_policy": "every step renders new iterations (new random streams) of the whole image; path state of the CTA-local "
                               "wavefront lives in shared memory, sample planes (16 B/sample) stream to HBM; no explicit flush",
                  "parallelism": f"tile-sharded x{world}, one NCCL reduce per step inside libb200pt" if world > 1 else "single GPU"}
        line = {"metric": "Msamples/s", "value": res["value"], "unit": "Msamples/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "device_ms_per_step": res["dev_ms"] / a.steps,
                "e2e": {"value": res["e2e"], "unit": "Msamples/s", "h2d_bytes_per_step": 104, "d2h_bytes_per_step": res["npix"] * 12},
                "gpu_launches": int(res["launches"]), "rays_per_sample": res["rays"] / res["samples"], "clocks": res["clocks"],
                "roofline": roofline, "cpu_baseline": cb, "reference_cuda": ref_cuda, "extra": extra}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
