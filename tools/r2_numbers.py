#!/usr/bin/env python
"""Markdown summary of a bench line (profiles/bench_r2.json by default) for DESIGN.md / README.md."""
import json, sys
f = sys.argv[1] if len(sys.argv) > 1 else "profiles/bench_r2.json"
f_path = f
d = json.loads(open(f).read().strip().splitlines()[-1])
ref = None
try:
    ref = json.loads(open(f.replace("bench_", "bench_reference_")).read().strip().splitlines()[-1])
except Exception:
    pass


import os
KFILE = {"C3": "agg_kernel", "C2": "scan_kernel", "C4": "topn_kernel", "C5": "checksum_kernel"}


def traffic_ratio(name, rf, rows):
    """DRAM bytes per entry of the kernel's ncu capture / algorithmic bytes per entry of the bench workload."""
    f = os.path.join(os.path.dirname(f_path), KFILE.get(name, "") + "_r2_traffic.json")
    if not os.path.exists(f):
        return None
    return json.load(open(f))["dram_bytes_per_entry"] / (rf["algorithmic_bytes_per_step"] / rows)


def row(r, rows):
    rf = r["roofline"]
    name = r.get('config', {}).get('workload', r.get('workload', '')).split(':')[0]
    tr = traffic_ratio(name, rf, rows)
    t = tr * rf["algorithmic_bytes_per_step"] if tr else None
    return (f"| {r.get('config', {}).get('workload', r.get('workload', '')).split(':')[0]} | {rows:.0e} | {r['ms_per_step']:.2f} | {rows * d['n_gpus'] / (r['ms_per_step'] / 1e3):.3g} | "
            f"{rf['kernel_ms_per_step']:.2f} | {rf['achieved']:.0f} | **{rf['frac']:.3f}** | {(t / rf['algorithmic_bytes_per_step']):.2f}× |" if t else
            f"| {r.get('config', {}).get('workload', r.get('workload', '')).split(':')[0]} | {rows:.0e} | {r['ms_per_step']:.2f} | {rows * d['n_gpus'] / (r['ms_per_step'] / 1e3):.3g} | "
            f"{rf['kernel_ms_per_step']:.2f} | {rf['achieved']:.0f} | **{rf['frac']:.3f}** | n/a |") + f" {r.get('merge_ms', 0):.2f} |"


print("| workload | rows / GPU | ms / step | rows/s | unit-kernel ms | GB/s | frac of measured HBM peak | DRAM traffic / algorithmic | merge ms |")
print("|---|---|---|---|---|---|---|---|---|")
print(row(d, d["config"].get("rows_per_gpu", d.get("rows_per_gpu", 0)) or d["config"].get("rows", 0)))
for s in d.get("sub", []):
    print(row(s, s["rows_per_gpu"]))
e = d.get("e2e") or {}
print()
print(f"value {d['value']:.4g} {d['unit']}, n_gpus {d['n_gpus']}, clocks {d.get('clocks')}, gpu_launches {d.get('gpu_launches')}")
print(f"e2e cold {e.get('value', 0):.4g} rows/s ({e.get('ms_per_step', 0):.1f} ms, H2D {e.get('h2d_bytes_per_step', 0) / 1e9:.2f} GB, rows/GPU {e.get('rows_per_gpu')}), warm {e.get('warm', {}).get('value', 0):.4g} rows/s ({e.get('warm', {}).get('ms_per_step', 0):.2f} ms)")
if e.get("flat"):
    print(f"e2e source: {e.get('source')}; compressed / flat bytes {e.get('compressed_to_flat_bytes', 0):.3f}, device expansion {e.get('decode_ms_per_step', 0):.0f} ms per step (overlapped with H2D)")
    print(f"e2e from flat CF blocks: {e['flat']['value']:.4g} rows/s ({e['flat']['ms_per_step']:.1f} ms, H2D {e['flat']['h2d_bytes_per_step'] / 1e9:.2f} GB)")
print("cpu_baseline", d.get("cpu_baseline"))
if ref:
    print("reference arm", ref.get("value"), ref.get("unit"), ref.get("cpu_baseline"))
