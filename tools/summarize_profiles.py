#!/usr/bin/env python
"""gpurun_out/ (scratch) -> profiles/ (tracked): bench lines, launch list, and the headline numbers of the ncu --set full
capture of the dominant kernel, including the DRAM traffic figure bench.py reports as roofline.traffic."""
import csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r1"
ENTRIES = int(sys.argv[2]) if len(sys.argv) > 2 else 12_500_000  # entries covered by the captured launch (one 12.5M-entry block)
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
for f in (f"bench_{R}.json", f"bench_generic_kernel_{R}.json", f"bench_reference_{R}.json", f"launches_{R}.csv", f"other_configs_{R}.jsonl", f"scan_kernel_{R}.ncu-rep"):
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(P, f))
rep = os.path.join(P, f"scan_kernel_{R}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
m = {n: (v[i], u[i]) for i, n in enumerate(h)}
def num(name):
    x, unit = m[name]
    x = float(x.replace(",", ""))
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}.get(unit, 1)
    return x * scale
keep = ["sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
stalls = {n.split("issue_stalled_")[1].split("_per_issue")[0]: float(m[n][0]) for n in h if "average_warps_issue_stalled" in n and n.endswith(".ratio") and m[n][0]}
rd, wr, dur = num("dram__bytes_read.sum"), num("dram__bytes_write.sum"), num("gpu__time_duration.sum")
full = {"source": f"profiles/scan_kernel_{R}.ncu-rep (ncu --set full --import-source on --clock-control none -k 'regex:scan_kernel|b2_scan_jit' -s 8 -c 1; bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --chunk 12500000: one launch = one 12.5M-entry block)",
        "kernel": m["Kernel Name"][0] if "Kernel Name" in m else "scan_kernel<PM_SCAN>", "duration_us_under_ncu": dur * 1e6,
        "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_GBps_under_ncu": (rd + wr) / dur / 1e9,
        "metrics": {k: m[k][0] + (" " + m[k][1] if m[k][1] else "") for k in keep if k in m},
        "stall_cycles_per_issued_instruction": dict(sorted(stalls.items(), key=lambda kv: -kv[1]))}
json.dump(full, open(os.path.join(P, f"scan_kernel_{R}_ncu_full.json"), "w"), indent=1)
traffic = {"source": full["source"], "kernel": full["kernel"], "entries_in_launch": ENTRIES, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "dram_bytes_per_entry": (rd + wr) / ENTRIES,
           "note": "per-launch DRAM traffic of the dominant kernel; bench.py scales it to its own launch size for roofline.traffic"}
json.dump(traffic, open(os.path.join(P, f"scan_kernel_{R}_traffic.json"), "w"), indent=1)
# launch list: share of the step per kernel
lr = [r for r in csv.reader(l for l in open(os.path.join(P, f"launches_{R}.csv")) if not l.startswith("==")) if r]
hi = lr[0]
ki, vi = hi.index("Kernel Name"), hi.index("Metric Value")
tot = {}
for r in lr[1:]:
    try:
        tot[r[ki].split("(")[0]] = tot.get(r[ki].split("(")[0], 0) + float(r[vi].replace(",", ""))
    except (ValueError, IndexError):
        pass
setup = {k: t for k, t in tot.items() if "gen_" in k or "cub::" in k}  # synthetic table generator: outside every timed region
tot = {k: t for k, t in tot.items() if k not in setup}
s = sum(tot.values())
share = {k: {"ns": t, "share_of_step": t / s} for k, t in sorted(tot.items(), key=lambda kv: -kv[1])}
share["(setup, not part of a step) generator kernels"] = {"ns": sum(setup.values())}
json.dump({"source": f"profiles/launches_{R}.csv (ncu --metrics gpu__time_duration.sum, bench.py --steps 2 --warmup 3 --no-e2e --no-cpu, generator kernels included)",
           "per_kernel": share}, open(os.path.join(P, f"launches_{R}_share.json"), "w"), indent=1)
print(json.dumps(traffic)); print(json.dumps(full["metrics"])); print({k: round(x.get("share_of_step", 0), 4) for k, x in share.items()})
