#!/usr/bin/env python
"""Single-GPU throughput of the other BASELINE configs (not the bench.py headline): C3 hash-agg, C4 TopN, C5 checksum.
Prints one JSON line per workload: rows/s, kernel-only GB/s against the measured HBM peak.  Inputs resident in HBM."""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def gen(ffi, n_rows, n_blocks, n_cols, lo, rng, nulls=None, seed=0x525C682A2F7CE3DB, table=1000, dirty=None):
    L = ffi.lib()
    gens, blks = [], []
    per = (n_rows + n_blocks - 1) // n_blocks
    h, left = 0, n_rows
    keep = []
    while left > 0:
        n = min(per, left)
        s = ffi.GenSpec()
        s.table_id, s.first_handle, s.n_rows, s.n_cols, s.row_format, s.seed = table, h, n, n_cols, 2, seed
        a, b = (C.c_int64 * n_cols)(*lo), (C.c_uint64 * n_cols)(*rng)
        s.col_lo, s.col_range = a, b
        keep += [a, b]
        if nulls:
            c = (C.c_uint32 * n_cols)(*nulls)
            s.null_per_million = c
            keep.append(c)
        s.commit_ts, s.newer_ts = 20, 5000
        if dirty:
            s.extra_versions_per_million, s.delete_per_million, s.lock_rec_per_million = dirty
        g, blk = C.c_void_p(), ffi.GenBlock()
        assert L.b2_gen_create(0, C.byref(s), C.byref(g), C.byref(blk)) == 0, L.b2_last_error_message()
        gens.append(g); blks.append(blk)
        h += n; left -= n
    return gens, blks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=200_000_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--only", default="", help="comma list of: c3,c4 (default all)")
    args = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    import bench
    from tikv_b200 import ffi
    from tikv_b200.executor import BatchExecutor, checksum
    from tikv_b200.plan import ColumnDef, Plan, col, const_int
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    rng_tab = bench.table_range()

    def run(name, plan, src, in_bytes, rows):
        ffi.lib().b2_plan_prepare(C.byref(plan.c), 0)  # prepared plan: the specialised kernel is compiled before the timed runs
        best = None
        for _ in range(args.steps + 1):
            t0 = time.perf_counter()
            with BatchExecutor(plan, rng_tab, src, output=ffi.LOC_DEVICE) as ex:
                rc, b = ex.next_batch_raw(1 << 40)
                assert rc == 0, ex.last_error().message
                st = ex.collect_exec_stats()
                n_out = b.n_rows
            dt = time.perf_counter() - t0
            best = (dt, st.kernel_time_ns / 1e9, n_out) if best is None or dt < best[0] else best
        dt, kt, n_out = best
        print(json.dumps({"workload": name, "rows": rows, "rows_per_s": rows / dt, "ms": dt * 1e3, "kernel_ms": kt * 1e3,
                          "kernel_GBps": in_bytes / kt / 1e9, "frac_of_measured_hbm": in_bytes / kt / 1e9 / peak, "out_rows": n_out}))

    only = set(x for x in args.only.split(",") if x)
    # C3: GROUP BY i32 key SUM(i64); key uniform in [0, G)
    for G in ((2, 1024, 1 << 20) if (not only or "c3" in only) else ()):
        gens, blks = gen(ffi, args.rows, 8, 2, [0, -(1 << 40)], [G, 1 << 41])
        src = bench.Source(ffi, [b.block for b in blks], ffi.LOC_DEVICE, 0)
        in_bytes = sum(b.key_bytes + b.val_bytes + 8 * b.block.n for b in blks)
        cols = [ColumnDef(100, pk_handle=True), ColumnDef(1, tp=ffi.TP_LONG), ColumnDef(2)]
        plan = Plan().table_scan(1000, cols).aggregation([("sum", col(2))], group_by=[col(1, tp=ffi.TP_LONG)]).build()
        run(f"C3 FastHashAgg GROUP BY i32 (G={G}) SUM(i64)", plan, src, in_bytes, args.rows)
        if G == 1024:
            plan = Plan().table_scan(1000, cols).aggregation([("count", const_int(1)), ("sum", col(2))]).build()
            run("C1-style SimpleAgg COUNT(*) + SUM(i64)", plan, src, in_bytes, args.rows)
            # C5: checksum over the same KVs
            best = None
            for _ in range(args.steps):
                t0 = time.perf_counter()
                rc, res, msg = checksum(rng_tab, src)
                dt = time.perf_counter() - t0
                assert rc == 0, msg
                best = dt if best is None or dt < best else best
            print(json.dumps({"workload": "C5 checksum CRC64-XZ", "kvs": res[1], "bytes": res[2], "kvs_per_s": res[1] / best, "ms": best * 1e3,
                              "GBps": (res[2] + 8 * res[1]) / best / 1e9, "frac_of_measured_hbm": (res[2] + 8 * res[1]) / best / 1e9 / peak}))
        for g in gens:
            ffi.lib().b2_gen_destroy(g)
    # BatchSlowHashAggregation: GROUP BY two i32 keys, SUM(i64); G = G1 * G2 groups
    for G1, G2 in (((2, 2), (16, 4), (1024, 16), (1024, 1024)) if (not only or "c3m" in only) else ()):
        gens, blks = gen(ffi, args.rows, 8, 3, [0, 0, -(1 << 40)], [G1, G2, 1 << 41])
        src = bench.Source(ffi, [b.block for b in blks], ffi.LOC_DEVICE, 0)
        in_bytes = sum(b.key_bytes + b.val_bytes + 8 * b.block.n for b in blks)
        cols = [ColumnDef(100, pk_handle=True), ColumnDef(1, tp=ffi.TP_LONG), ColumnDef(2, tp=ffi.TP_LONG), ColumnDef(3)]
        plan = Plan().table_scan(1000, cols).aggregation([("sum", col(3))], group_by=[col(1, tp=ffi.TP_LONG), col(2, tp=ffi.TP_LONG)]).build()
        run(f"SlowHashAgg GROUP BY (i32, i32) (G={G1}x{G2}) SUM(i64)", plan, src, in_bytes, args.rows)
        for g in gens:
            ffi.lib().b2_gen_destroy(g)
    if not only or "dirty" in only:
        # C2 on a "dirty" table: 30 % of the keys carry extra (older / newer-than-read_ts) versions, 5 % are
        # deleted, 5 % have a Lock/Rollback record on top: version runs, skipped records, met_newer_ts_data
        gens, blks = gen(ffi, args.rows, 8, 8, [0] * 8, [0] * 8, dirty=(300000, 50000, 50000))
        src = bench.Source(ffi, [b.block for b in blks], ffi.LOC_DEVICE, 0)
        in_bytes = sum(b.key_bytes + b.val_bytes + 8 * b.block.n for b in blks)
        n_entries = sum(b.block.n for b in blks)
        run(f"C2 scan + selection on a dirty table ({n_entries} CF_WRITE entries for {args.rows} keys)", bench.build_plan("c2"), src, in_bytes, args.rows)
        for g in gens:
            ffi.lib().b2_gen_destroy(g)
    if only and "c4" not in only:
        return
    # C4: TopN ORDER BY c0 DESC, c1 ASC LIMIT 1000, second column 1 % NULL
    gens, blks = gen(ffi, args.rows, 8, 2, [0, 0], [0, 0], nulls=[0, 10000])
    src = bench.Source(ffi, [b.block for b in blks], ffi.LOC_DEVICE, 0)
    in_bytes = sum(b.key_bytes + b.val_bytes + 8 * b.block.n for b in blks)
    cols = [ColumnDef(100, pk_handle=True), ColumnDef(1), ColumnDef(2)]
    plan = Plan().table_scan(1000, cols).topn([(col(1), True), (col(2), False)], 1000).build()
    run("C4 TopN ORDER BY 2 x i64 LIMIT 1000", plan, src, in_bytes, args.rows)
    for g in gens:
        ffi.lib().b2_gen_destroy(g)


if __name__ == "__main__":
    main()
