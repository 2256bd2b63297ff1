#!/usr/bin/env python
"""gpurun_out/ (scratch) -> profiles/ (tracked), round 2: bench lines, launch list with each kernel's share of the step,
and per dominant kernel (C3 aggregation, C2 scan, C4 TopN, C5 checksum) the headline numbers of its `ncu --set full`
capture, including the DRAM traffic per entry that bench.py reports as roofline.traffic."""
import csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r2"
ENTRIES = int(sys.argv[2]) if len(sys.argv) > 2 else 12_500_000  # entries covered by each captured launch (one block)
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KERNELS = {"agg_kernel": "C3 lean kernel (fast_body<PM_AGG>, plan-specialised): scan + selection + hash aggregation",
           "scan_kernel": "C2 scan kernel (scan_body<PM_SCAN>, plan-specialised): scan + selection + ordered compaction",
           "topn_kernel": "C4 lean kernel (fast_body<PM_TOPN>, plan-specialised)",
           "checksum_kernel": "C5 lean kernel (fast_body<PM_CHECKSUM>)",
           "sst_expand": "block reader, expansion pass (sst_expand_kernel, one warp per restart interval): 2e7 C3 entries of RocksDB data blocks -> flat block"}
ENTRIES_OF = {"sst_expand": 20_000_000}
for f in [f"bench_{R}.json", f"bench_reference_{R}.json", f"launches_{R}.csv", f"sst_probe_{R}.log"] + [f"{k}_{R}.ncu-rep" for k in KERNELS]:
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(P, f))
keep = ["sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
for k, what in KERNELS.items():
    rep = os.path.join(P, f"{k}_{R}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    ENTRIES = ENTRIES_OF.get(k, int(sys.argv[2]) if len(sys.argv) > 2 else 12_500_000)
    h, u, v = rows[0], rows[1], rows[2]
    m = {n: (v[i], u[i]) for i, n in enumerate(h)}

    def num(name):
        x, unit = m[name]
        return float(x.replace(",", "")) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}.get(unit, 1)
    stalls = {n.split("issue_stalled_")[1].split("_per_issue")[0]: float(m[n][0]) for n in h if "average_warps_issue_stalled" in n and n.endswith(".ratio") and m[n][0]}
    rd, wr, dur = num("dram__bytes_read.sum"), num("dram__bytes_write.sum"), num("gpu__time_duration.sum")
    src = f"profiles/{k}_{R}.ncu-rep (tools/refresh_profiles_r2.sh: ncu --set full --import-source on --clock-control none, one launch = {ENTRIES} entries)"
    inst = float(m["smsp__inst_executed.sum"][0].replace(",", ""))
    full = {"source": src, "what": what, "kernel": m["Kernel Name"][0] if "Kernel Name" in m else k, "duration_us_under_ncu": dur * 1e6,
            "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_GBps_under_ncu": (rd + wr) / dur / 1e9,
            "warp_instructions_per_32_entries": inst / (ENTRIES / 32),
            "metrics": {x: m[x][0] + (" " + m[x][1] if m[x][1] else "") for x in keep if x in m},
            "stall_cycles_per_issued_instruction": dict(sorted(stalls.items(), key=lambda kv: -kv[1]))}
    json.dump(full, open(os.path.join(P, f"{k}_{R}_ncu_full.json"), "w"), indent=1)
    json.dump({"source": src, "kernel": full["kernel"], "entries_in_launch": ENTRIES, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_entry": (rd + wr) / ENTRIES,
               "note": "per-launch DRAM traffic of the dominant kernel; bench.py scales it to its own launch size for roofline.traffic"},
              open(os.path.join(P, f"{k}_{R}_traffic.json"), "w"), indent=1)
    print(k, round(dur * 1e6, 1), "us", round((rd + wr) / ENTRIES, 1), "B/entry", round(inst / (ENTRIES / 32)), "inst/32 entries")
lp = os.path.join(P, f"launches_{R}.csv")
if os.path.exists(lp):
    lr = [r for r in csv.reader(l for l in open(lp) if not l.startswith("==")) if r]
    hi = lr[0]
    ki, vi, ui = hi.index("Kernel Name"), hi.index("Metric Value"), hi.index("Metric Unit")
    tot = {}
    for r in lr[1:]:
        try:
            t = float(r[vi].replace(",", "")) * {"us": 1e3, "ms": 1e6, "ns": 1}.get(r[ui], 1)
            tot[r[ki].split("(")[0]] = tot.get(r[ki].split("(")[0], 0) + t
        except (ValueError, IndexError):
            pass
    setup = {k: t for k, t in tot.items() if "gen_" in k or "cub::" in k}  # synthetic table generator: outside every timed region
    tot = {k: t for k, t in tot.items() if k not in setup}
    s = sum(tot.values())
    share = {k: {"ns": t, "share_of_gpu_time": t / s} for k, t in sorted(tot.items(), key=lambda kv: -kv[1])}
    share["(setup, not part of a step) generator kernels"] = {"ns": sum(setup.values())}
    json.dump({"source": f"profiles/launches_{R}.csv (ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity: "
                         "the headline C3 steps followed by the C2 / C5 / C4 sub-records, generator kernels listed apart)",
               "per_kernel": share}, open(os.path.join(P, f"launches_{R}_share.json"), "w"), indent=1)
    print({k: round(x.get("share_of_gpu_time", 0), 4) for k, x in list(share.items())[:8]})
