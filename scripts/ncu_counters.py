#!/usr/bin/env python
"""GPU box: per-workload hardware counters of the wavefront kernels, measured with ncu, written as JSON for bench.py.

    python scripts/ncu_counters.py --out profiles/r02_counters.json [--workloads c2,c4]

For every workload one small render (same scene, same image size as bench.py's workload, a few spp) runs under
`ncu --metrics ... --clock-control none`; the per-launch rows are summed per kernel and normalised per SAMPLE:
warp instructions, thread instructions (-> active lanes per instruction), DRAM bytes read + written, kernel time.
bench.py multiplies these per-sample figures with the samples of its timed region (roofline.issue / roofline.traffic)
instead of carrying constants in its source.  Numbers taken under the profiler are never bench values: only the
per-sample COUNTS are used (they do not depend on timing), and the kernels' time SHARES."""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "dram__bytes_read.sum",
           "dram__bytes_write.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum", "lts__t_sector_hit_rate.pct"]
WORK = {                         # scene name for scripts/compare_ref.make, size, spp of the profiled render
    "c2": ("cornell", 1024, 4),
    "c1": ("cornell4", 256, 16),
    "c3": ("veach", 768, 4),
    "c4": ("tris1000000", 2048, 1),
    "c5": ("vol", 512, 8),
    "smoke": ("smoke", 1024, 2),
    "shipped512": ("shipped", 512, 4),
    "zoo": ("zoo", 512, 8),
}


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def run(workload):
    scene, size, spp = WORK[workload]
    cmd = ["ncu", "--metrics", ",".join(METRICS), "--clock-control", "none", "--csv", "-k", "regex:k_(wave|shade|trace|resolve|het)",
           sys.executable, os.path.join(ROOT, "scripts", "compare_ref.py"), "--scene", scene, "--size", str(size), "--spp", str(spp), "--no-ref",
           "--no-warm"]
    out = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT).stdout
    lines = out.splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith('"ID"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    res = next((json.loads(ln[7:]) for ln in lines if ln.startswith("RESULT ")), {})
    samples = float(res.get("w", 0)) * float(res.get("h", 0)) * float(res.get("spp", 0))
    per = {}
    for r in rows:
        k = r["Kernel Name"].split("(")[0]
        e = per.setdefault(k, {"launches": set(), **{m: 0.0 for m in METRICS}})
        e["launches"].add(r["ID"])
        v = num(r["Metric Value"])
        unit = r.get("Metric Unit", "")
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0, "second": 1e3}.get(unit, 1e-6)
        if m.startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        if "pct" in m:
            e[m] = max(e[m], v)          # percentages: report the largest launch's (not summable)
        else:
            e[m] += v
    kernels = {}
    tot_ms = sum(e["gpu__time_duration.sum"] for e in per.values())
    for k, e in per.items():
        wi = e["smsp__inst_executed.sum"]; ti = e["smsp__thread_inst_executed.sum"]
        kernels[k] = {"launches": len(e["launches"]), "time_share": e["gpu__time_duration.sum"] / max(tot_ms, 1e-9),
                      "ms_under_ncu": e["gpu__time_duration.sum"],
                      "warp_inst_per_sample": wi / samples, "active_lanes_per_inst": ti / max(wi, 1.0),
                      "dram_bytes_per_sample": (e["dram__bytes_read.sum"] + e["dram__bytes_write.sum"]) / samples,
                      "fma_pipe_inst_per_sample": e["smsp__inst_executed_pipe_fma.sum"] / samples,
                      "alu_pipe_inst_per_sample": e["smsp__inst_executed_pipe_alu.sum"] / samples,
                      "issue_active_pct": e["smsp__issue_active.avg.pct_of_peak_sustained_active"],
                      "warps_active_pct": e["sm__warps_active.avg.pct_of_peak_sustained_active"],
                      "l2_hit_pct": e["lts__t_sector_hit_rate.pct"]}
    wi = sum(e["smsp__inst_executed.sum"] for e in per.values()); ti = sum(e["smsp__thread_inst_executed.sum"] for e in per.values())
    dr = sum(e["dram__bytes_read.sum"] + e["dram__bytes_write.sum"] for e in per.values())
    top = max(kernels, key=lambda k: kernels[k]["time_share"]) if kernels else None
    return {"scene": scene, "size": size, "spp": spp, "samples": samples, "dominant_kernel": top,
            "warp_inst_per_sample": wi / samples, "active_lanes_per_inst": ti / max(wi, 1.0), "dram_bytes_per_sample": dr / samples,
            "kernels": kernels, "command": " ".join(cmd[:9]) + " ... compare_ref.py --scene %s --size %d --spp %d" % (scene, size, spp)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "counters.json"))
    ap.add_argument("--workloads", default="c2,c4")
    a = ap.parse_args()
    data = {}
    if os.path.exists(a.out):
        data = json.load(open(a.out))
    for w in a.workloads.split(","):
        data[w] = run(w)
        print(w, json.dumps({k: v for k, v in data[w].items() if k != "kernels"}), flush=True)
        for k, v in data[w]["kernels"].items():
            print("   ", k, json.dumps(v), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(data, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
