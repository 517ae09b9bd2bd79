#!/usr/bin/env python
"""Condense an `ncu --set full` report (exported with `ncu -i X.ncu-rep --page raw --csv`) into the per-kernel figures
quoted in profiles/: duration, registers, DRAM bytes, pipe utilisation, issue activity and the top stall reasons.
usage: ncu_key_metrics.py raw.csv out.json > out.txt"""
import csv, json, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
def f(x):
    try: return float(x.replace(",", ""))
    except Exception: return None
KEYS = {
    "duration_us": "gpu__time_duration.sum",
    "registers": "launch__registers_per_thread",
    "dram_bytes_read": "dram__bytes_read.sum",
    "dram_bytes_write": "dram__bytes_write.sum",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "fp64_pipe_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "lsu_pipe_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "inst_executed": "smsp__inst_executed.sum",
    "grid": "launch__grid_size", "block": "launch__block_size", "smem_dyn": "launch__shared_mem_per_block_dynamic",
}
units = dict(zip(hdr, rows[1]))
out, seen = {}, set()
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("unnamed>::", "").strip()
    if name in seen:
        if not name.startswith("eval_tmem_kernel") or name.replace("eval_tmem_kernel", "eval_tmem_kernel[imu items]") in seen: continue
        name = name.replace("eval_tmem_kernel", "eval_tmem_kernel[imu items]")   # the evaluation kernel goes out twice: vision items, then IMU items
    seen.add(name)
    e = {}
    for k, m in KEYS.items():
        if m in d and f(d[m]) is not None:
            v = f(d[m]); u = units.get(m, "")
            if k == "duration_us" and u in ("ns", "nsecond"): v /= 1000.0
            if k.startswith("dram_bytes") and u.lower().startswith("k"): v *= 1e3
            if k.startswith("dram_bytes") and u.lower().startswith("m"): v *= 1e6
            e[k] = v
    stalls = sorted(((f(d[k]) or 0.0, k.split("issue_stalled_")[1].split("_per_issue")[0]) for k in hdr
                     if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k), reverse=True)
    e["top_stalls_per_issue"] = {n: round(v, 3) for v, n in stalls[:5]}
    out[name] = e
    print(f"== {name}")
    for k, v in e.items(): print(f"   {k:22s} {v}")
json.dump(out, open(sys.argv[2], "w"), indent=1)
