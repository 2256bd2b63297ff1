"""Static look at the plan-specialised kernel of a bench plan, without a GPU: NVRTC -> cubin -> registers / spills, SASS
instruction count, and SASS instructions per source function (nvdisasm line info).  Straight-line hot paths (the clean-entry
front end) can be budgeted this way before any GPU time is spent.

Usage: python tools/jit_sass.py c2|c3|c3f|c4 [out.cubin]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from cuda.bindings import nvrtc  # noqa: E402

import jit_compile_check as J  # noqa: E402
from tikv_b200 import ffi  # noqa: E402
from tikv_b200.plan import ColumnDef, Plan, col, const_int, lt  # noqa: E402


def plan_of(name):
    if name == "c2":
        cols = [ColumnDef(100, pk_handle=True)] + [ColumnDef(i + 1) for i in range(8)]
        return Plan().table_scan(1000, cols).selection(lt(col(1), const_int(0))).build(output_offsets=list(range(1, 9)))
    cols = [ColumnDef(100, pk_handle=True), ColumnDef(1, tp=ffi.TP_LONG), ColumnDef(2)]
    if name == "c3":
        return Plan().table_scan(1000, cols).aggregation([("sum", col(2))], group_by=[col(1, tp=ffi.TP_LONG)]).build()
    if name == "c3f":
        return Plan().table_scan(1000, cols).selection(lt(col(2), const_int(0))).aggregation([("sum", col(2))], group_by=[col(1, tp=ffi.TP_LONG)]).build()
    if name == "c4":
        cols = [ColumnDef(100, pk_handle=True), ColumnDef(1), ColumnDef(2)]
        return Plan().table_scan(1000, cols).topn([(col(1), True), (col(2), False)], 1000).build()
    raise SystemExit("unknown plan " + name)


def compile_cubin(plan):
    lit = J.literal_of(plan)
    mode = J.mode_of(lit)
    if not os.environ.get("B2_KEEP_V1"):  # the engine clears fast_v1 when the data is row format v2 (it samples the first row)
        lit, n = re.subn(r"(u,\{-?\d+(?:,-?\d+){7}\},\d+,\d+u),1,", r"\1,0,", lit, count=1)
    src = (os.environ.get("B2_SRC_DEFS", "#define B2_NO_IDX 1\n") + "#define B2_NVRTC 1\n#define B2_JIT_PLAN 1\n#include \"fast_kernel.cuh\"\nnamespace b2 { __constant__ const DevPlan kJitPlan =\n" + lit + ";\n}\n"
           "extern \"C\" __global__ void __launch_bounds__(b2::TILE + 64, 2) b2_scan_jit(const __grid_constant__ b2::ScanArgs A) {\n"
           "  b2::scan_body<" + str(mode) + ">(b2::kJitPlan, A);\n}\n")
    if mode in (1, 2) and os.environ.get("B2_FAST", "1") == "1":
        src += ("extern \"C\" __global__ void __launch_bounds__(b2::FK_THREADS, " + os.environ.get("B2_FAST_MINB", "2") + ") b2_fast_jit(const __grid_constant__ b2::ScanArgs A) {\n"
                "  b2::fast_body<" + str(mode) + ">(b2::kJitPlan, A);\n}\n")
    err, prog = nvrtc.nvrtcCreateProgram(src.encode(), b"b2_scan_jit.cu", 0, [], [])
    opts = [b"--gpu-architecture=sm_100a", b"--std=c++17", b"-lineinfo", b"-DB2_NVRTC=1", b"-default-device",
            ("-I" + os.path.join(ROOT, "tikv_b200", "csrc")).encode(), b"-I/usr/local/cuda/include", b"--ptxas-options=-v", b"-DB2_EXT_SIGS=0"] + [o.encode() for o in os.environ.get("B2_JIT_EXTRA", "").split()]
    (rc,) = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    _, n = nvrtc.nvrtcGetProgramLogSize(prog)
    log = b" " * n
    nvrtc.nvrtcGetProgramLog(prog, log)
    if rc != nvrtc.nvrtcResult.NVRTC_SUCCESS:
        raise SystemExit(log.decode(errors="replace")[:3000])
    _, n = nvrtc.nvrtcGetCUBINSize(prog)
    cubin = b" " * n
    nvrtc.nvrtcGetCUBIN(prog, cubin)
    return cubin, log.decode(errors="replace")


def main():
    name = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else f"/tmp/jit/{name}.cubin"
    cubin, log = compile_cubin(plan_of(name))
    open(out, "wb").write(cubin)
    print(" | ".join(l.strip() for l in log.splitlines() if "registers" in l or "spill" in l))
    dis = subprocess.run(["nvdisasm", "-g", "-c", out], capture_output=True, text=True).stdout
    # nvdisasm -g interleaves '//## File "...", line N' markers with instructions
    cur = None
    per_line = collections.Counter()
    ops = collections.Counter()
    total = 0
    want_fn = os.environ.get("FN", "b2_fast_jit" if "b2_fast_jit" in dis else "b2_scan_jit")
    fn = None
    per_fn = collections.Counter()
    for l in dis.splitlines():
        mf = re.match(r"\s*\.text\.(\w+):", l) or re.match(r"^(b2_\w+):", l.strip())
        if mf:
            fn = mf.group(1)
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", l)
        if m:
            per_fn[fn] += 1
        if fn != want_fn:
            mm = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if mm:
                cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", l)
        if m:
            total += 1
            per_line[cur] += 1
            ops[m.group(2).split(".")[0]] += 1
    print("per kernel:", dict(per_fn), "| buckets below are for", want_fn)
    print("SASS instructions:", total)
    print("ops:", ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))
    # per function of b2_device.h / 20-line bucket of scan_kernel.cuh
    src = open(os.path.join(ROOT, "tikv_b200", "csrc", "b2_device.h")).read().split("\n")
    funcs = []
    for i, l in enumerate(src, 1):
        m = re.match(r"^(?:template <[^>]*>\s*)?B2_HD\s+[\w:<>\* ]+?\s+(\w+)\(", l)
        if m:
            funcs.append((i, m.group(1)))

    def func_of(line):
        name = "?"
        for i, n in funcs:
            if i <= line:
                name = n
            else:
                break
        return name
    agg = collections.Counter()
    for (f, ln), c in ((k, v) for k, v in per_line.items() if k):
        agg[("dev:" + func_of(ln)) if f == "b2_device.h" else f"{f}:{ln // 20 * 20}"] += c
    for k, v in agg.most_common(int(os.environ.get("TOPK", "40"))):
        print(f"  {k:40s} {v}")


if __name__ == "__main__":
    main()
