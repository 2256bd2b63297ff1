#!/usr/bin/env python
"""bench.py — coprocessor rows/sec on B200 (BASELINE.json metric), one JSON line on rank 0.

Workload (BASELINE.json configs[1], SURVEY.md §8(d) "C2"): BatchTableScan + BatchSelection `col0 < k` over a
region-sharded synthetic table of 1e8 rows x 8 i64 columns (row format v2, one version per key), k = 0 (50 %).
A "step" is one pass of the hot path over the whole table:
  value : inputs already resident in HBM, results left in HBM (device time, CUDA events on the launch stream)
  e2e   : same request through the C ABI with HOST buffers (pinned): H2D of every block + D2H of the selected
          columns inside the timed region
  --impl reference : the CPU oracle (C++ restatement of the reference algorithm; the Rust reference cannot be
          built in this image) on all host cores, one region task per thread, on a bounded sample of the workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

TABLE_ID = 1000
SEED = 0x525C682A2F7CE3DB  # tests/benches/coprocessor_executors/util/fixture.rs:26
READ_TS = 1000
N_COLS = 8
METRIC = "coprocessor rows/sec (scan+filter+hash-agg) at 1/2/4/8 B200 vs host-CPU ref"


def build_plan():
    from tikv_b200.plan import ColumnDef, Plan, col, const_int, lt
    columns = [ColumnDef(100, pk_handle=True)] + [ColumnDef(i + 1) for i in range(N_COLS)]
    plan = Plan().table_scan(TABLE_ID, columns).selection(lt(col(1), const_int(0))).build(output_offsets=list(range(1, N_COLS + 1)))
    return plan


def table_range():
    import struct
    pre = b"t" + struct.pack(">Q", TABLE_ID ^ (1 << 63))
    return [(pre + b"_r", pre + b"_s")]


def gen_blocks(ffi, device, n_rows, n_blocks, first_handle=0, row_format=2):
    """Generate the table on the device as `n_blocks` CF_WRITE blocks.  Returns (gens, GenBlock list)."""
    L = ffi.lib()
    gens, blks = [], []
    per = (n_rows + n_blocks - 1) // n_blocks
    h = first_handle
    left = n_rows
    while left > 0:
        n = min(per, left)
        spec = ffi.GenSpec()
        spec.table_id, spec.first_handle, spec.n_rows, spec.n_cols, spec.row_format, spec.seed = TABLE_ID, h, n, N_COLS, row_format, SEED
        spec.commit_ts, spec.newer_ts = 20, 5000
        g, blk = C.c_void_p(), ffi.GenBlock()
        rc = L.b2_gen_create(device, C.byref(spec), C.byref(g), C.byref(blk))
        if rc != 0:
            raise RuntimeError("b2_gen_create: " + L.b2_last_error_message().decode())
        gens.append(g)
        blks.append(blk)
        h += n
        left -= n
    return gens, blks


class Source:
    def __init__(self, ffi, blocks, location, device):
        self.arr = (ffi.CfBlock * len(blocks))(*blocks)
        s = ffi.RegionSource()
        s.location, s.device, s.write, s.n_write = location, device, self.arr, len(blocks)
        s.read_ts, s.isolation_level, s.check_has_newer_ts_data = READ_TS, ffi.ISO_SI, 1
        self.c = s


def blocks_to_pinned_host(ffi, device, blks):
    """D2H copy of the generated blocks into pinned host buffers (setup, outside every timed region)."""
    L = ffi.lib()
    out, keep = [], []
    for b in blks:
        n = b.block.n
        sizes = [((b.key_bytes + 31) // 16) * 16, 4 * (n + 1), ((b.val_bytes + 31) // 16) * 16, 4 * (n + 1)]
        srcs = [b.block.keys, b.block.key_offs, b.block.vals, b.block.val_offs]
        copy = [b.key_bytes, 4 * (n + 1), b.val_bytes, 4 * (n + 1)]
        ptrs = []
        for sz, src, cb in zip(sizes, srcs, copy):
            p = L.b2_host_alloc_pinned(sz)
            if not p:
                raise RuntimeError("pinned host allocation failed")
            if L.b2_copy_to_host(device, p, src, cb) != 0:
                raise RuntimeError("D2H copy failed")
            ptrs.append(p)
            keep.append(p)
        hb = ffi.CfBlock()
        hb.keys, hb.key_offs, hb.vals, hb.val_offs, hb.n = ptrs[0], ptrs[1], ptrs[2], ptrs[3], n
        out.append(hb)
    return out, keep


def run_request(ffi, plan, ranges, src, out_loc, chunk, stream=0):
    """One step: open the executor, pull batches until drained.  Returns (rows_out, stats)."""
    from tikv_b200.executor import BatchExecutor
    rows = 0
    with BatchExecutor(plan, ranges, src, output=out_loc, stream=stream) as ex:
        while True:
            rc, b = ex.next_batch_raw(chunk)
            if rc != 0:
                raise RuntimeError("next_batch failed: " + ex.last_error().message.decode())
            rows += b.n_rows
            if b.is_drained != ffi.DRAIN_REMAIN:
                break
        st = ex.collect_exec_stats()
    return rows, st


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def workload_name(row_format):
    return f"C2: BatchTableScan + BatchSelection(col0 < 0) on 1e8 rows x 8 i64 cols, row format v{row_format}, 1 version/key (BASELINE.json configs[1])"


def usable_cores():
    """Host threads this process may actually run at once: the affinity mask, capped by the cgroup CPU quota (a
    container can see 128 CPUs and be allowed 32 of them; oversubscribing would understate the CPU arm)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    quota = None
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:  # cgroup v2: "<quota|max> <period>"
            q, per = f.read().split()[:2]
            if q != "max":
                quota = int(q) / int(per)
    except (OSError, ValueError):
        try:
            with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f, open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as g:  # cgroup v1
                q, per = int(f.read()), int(g.read())
                if q > 0 and per > 0:
                    quota = q / per
        except (OSError, ValueError):
            pass
    if quota is not None:
        n = min(n, max(1, int(quota + 0.999)))
    return max(1, n)


def cpu_reference_run(ffi, device, sample_rows, threads, steps, warmup, row_format=2):
    """The CPU arm: oracle (C++ restatement of the reference BatchExecutor pipeline), one region task per thread."""
    import orc
    L = orc.lib()
    plan = build_plan()
    from tikv_b200.plan import key_ranges
    kr, keep = key_ranges(table_range())
    gens, blks = gen_blocks(ffi, device, sample_rows, 1, row_format=row_format)
    host_blocks, pinned = blocks_to_pinned_host(ffi, device, blks)
    for g in gens:
        ffi.lib().b2_gen_destroy(g)
    src = Source(ffi, host_blocks, ffi.LOC_HOST, device)
    srcs = (ffi.RegionSource * threads)(*[src.c for _ in range(threads)])
    scanned, status = C.c_uint64(), C.c_int()
    times = []
    rows_out = 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        rows_out = L.orc_dag_handle_parallel(C.byref(plan.c), kr, 1, srcs, threads, threads, C.byref(scanned), C.byref(status))
        dt = time.perf_counter() - t0
        if status.value != 0:
            raise RuntimeError(f"oracle failed with status {status.value}")
        if it >= warmup:
            times.append(dt)
    for p in pinned:
        ffi.lib().b2_host_free_pinned(p)
    total_rows = sample_rows * threads
    sec = sum(times) / len(times)
    return {"rows_per_s": total_rows / sec, "sec_per_step": sec, "rows": total_rows, "rows_out": int(rows_out)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=100_000_000, help="rows per GPU")
    ap.add_argument("--blocks", type=int, default=8, help="CF_WRITE blocks (regions) per GPU")
    ap.add_argument("--chunk", type=int, default=1 << 24, help="CF_WRITE entries per next_batch")
    ap.add_argument("--cpu-sample-rows", type=int, default=2_000_000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-jit", action="store_true", help="generic kernel only")
    ap.add_argument("--row-format", type=int, default=2, choices=[1, 2], help="TiDB row format of the synthetic table (BASELINE quotes v2)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = usable_cores()

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    from tikv_b200 import ffi
    ffi.lib()
    device = local_rank if world > 1 else 0

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(ffi, device, args.cpu_sample_rows, cores, args.steps, args.warmup, args.row_format)
        line = {
            "impl": "reference", "metric": METRIC, "value": r["rows_per_s"], "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["sec_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic",
            "config": {"workload": workload_name(args.row_format), "rows_per_step": r["rows"], "selectivity": 0.5,
                       "sample": f"each step = {cores} region tasks x {args.cpu_sample_rows} rows of that table (a bounded sample of the workload)"},
            "cpu_baseline": {"value": r["rows_per_s"], "unit": "rows/s", "cores": cores, "kind": "port",
                             "sample": f"{cores} region tasks x {args.cpu_sample_rows} rows, one task per thread (restated C++ CPU baseline, not the TiKV Rust binary)"},
            "e2e": {"value": r["rows_per_s"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return

    torch.cuda.set_device(device)
    plan = build_plan()
    ranges = table_range()
    t_setup = time.time()
    # prepared plan: the scan kernel specialised for this DAG is compiled once per process (NVRTC), like a prepared statement;
    # when run-time compilation is unavailable the generic kernel serves the plan
    prep_rc = 0 if args.no_jit else ffi.lib().b2_plan_prepare(C.byref(plan.c), device)
    kernel_kind = "generic (interpreted plan)" if args.no_jit or prep_rc != 0 else "plan-specialised (compiled at run time, cached per plan)"
    if args.no_jit:
        os.environ["B2_JIT"] = "off"
    gens, blks = gen_blocks(ffi, device, args.rows, args.blocks, first_handle=rank * args.rows, row_format=args.row_format)
    dev_src = Source(ffi, [b.block for b in blks], ffi.LOC_DEVICE, device)
    n_entries = sum(b.block.n for b in blks)
    in_bytes = sum(b.key_bytes + b.val_bytes + 8 * b.block.n for b in blks)
    stream = torch.cuda.Stream(device=device)

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- HBM-resident steps ----
    rows_out, st = 0, None
    for _ in range(args.warmup):
        rows_out, st = run_request(ffi, plan, ranges, dev_src, ffi.LOC_DEVICE, args.chunk, stream.cuda_stream)
    sampler = ClockSampler(device)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    kernel_ns, launches, jit_launches = 0, 0, 0
    for _ in range(args.steps):
        rows_out, st = run_request(ffi, plan, ranges, dev_src, ffi.LOC_DEVICE, args.chunk, stream.cuda_stream)
        kernel_ns += st.kernel_time_ns
        launches += st.kernel_launches
        jit_launches += st.jit_launches
    ev1.record(stream)
    barrier()
    if args.no_e2e:
        clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], device=f"cuda:{device}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = args.rows * world / (ms_step / 1e3)
    out_bytes = rows_out * (N_COLS * 8) + N_COLS * ((rows_out + 63) // 64) * 8
    kernel_s = kernel_ns / 1e9 / args.steps
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = (in_bytes + out_bytes) / kernel_s / 1e9 if kernel_s > 0 else 0.0
    # DRAM traffic per launch from the committed ncu --set full capture (bytes per entry x entries per launch)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "scan_kernel_r1_traffic.json")
    if os.path.exists(tpath) and launches:
        traffic = json.load(open(tpath))["dram_bytes_per_entry"] * n_entries * args.steps / launches

    # ---- end to end with host buffers ----
    e2e = None
    if not args.no_e2e:
        host_blocks, pinned = blocks_to_pinned_host(ffi, device, blks)
        host_src = Source(ffi, host_blocks, ffi.LOC_HOST, device)
        for _ in range(2):
            r_e2e, st_e = run_request(ffi, plan, ranges, host_src, ffi.LOC_HOST, args.chunk, stream.cuda_stream)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        k = max(1, min(args.steps, 3))
        for _ in range(k):
            r_e2e, st_e = run_request(ffi, plan, ranges, host_src, ffi.LOC_HOST, args.chunk, stream.cuda_stream)
        e1.record(stream)
        barrier()
        te = torch.tensor([e0.elapsed_time(e1)], device=f"cuda:{device}")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        assert r_e2e == rows_out, "host-buffer path and HBM path disagree on the result size"
        e2e = {"value": args.rows * world / (float(te.item()) / k / 1e3), "unit": "rows/s", "h2d_bytes_per_step": int(st_e.h2d_bytes),
               "d2h_bytes_per_step": int(st_e.d2h_bytes), "ms_per_step": float(te.item()) / k}
        clocks = sampler.stop()  # sampled across both timed regions (HBM-resident steps and end-to-end steps)
        for p in pinned:
            ffi.lib().b2_host_free_pinned(p)
    for g in gens:
        ffi.lib().b2_gen_destroy(g)

    cpu = None
    if rank == 0 and not args.no_cpu:
        r = cpu_reference_run(ffi, device, args.cpu_sample_rows, cores, 1, 1, args.row_format)
        cpu = {"value": r["rows_per_s"], "unit": "rows/s", "cores": cores, "kind": "port",
               "sample": f"{cores} region tasks x {args.cpu_sample_rows} rows of the same workload, one task per thread (restated C++ CPU baseline, not the TiKV Rust binary)"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": workload_name(args.row_format),
                   "rows_per_gpu": args.rows, "cf_write_entries_per_gpu": n_entries, "blocks_per_gpu": args.blocks, "entries_per_batch": args.chunk,
                   "selectivity": rows_out / max(1, args.rows), "parallelism": f"region-sharded x{world}, no data-path collective",
                   "l2": f"inputs {in_bytes / 1e9:.1f} GB per pass >> 126 MB L2 (no flush needed)", "setup_s": round(time.time() - t_setup, 1)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes_per_launch": (in_bytes + out_bytes) * args.steps / max(1, launches),
                     "kernel": "scan_kernel<PM_SCAN>", "kernel_build": kernel_kind, "algorithmic_bytes_per_step": in_bytes + out_bytes, "kernel_ms_per_step": kernel_s * 1e3, "peak_source": peak_src},
        "e2e": e2e, "cpu_baseline": cpu, "gpu_launches": int(launches), "jit_launches": int(jit_launches), "clocks": clocks,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
