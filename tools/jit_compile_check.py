"""Compile the plan-specialised kernel of a plan with NVRTC, without a GPU (what jit.cu does at run time, minus the load).

Usage: python tools/jit_compile_check.py            # compiles every plan of the parity scenarios that the device accepts
Prints ptxas-level facts (registers, spills) per plan so that a broken or spilling specialisation shows up before GPU time
is spent."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cuda.bindings import nvrtc  # noqa: E402

from tikv_b200 import ffi  # noqa: E402


def literal_of(plan):
    L = ffi.lib()
    L.b2_plan_literal.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    L.b2_plan_literal.restype = C.c_int64
    buf = C.create_string_buffer(1 << 20)
    n = L.b2_plan_literal(C.byref(plan.c), buf, len(buf))
    return None if n < 0 else buf.value.decode()


def mode_of(lit):
    f = lit.strip("{").split(",")
    return int(f[0])


def compile_plan(plan, name):
    lit = literal_of(plan)
    if lit is None:
        return name, "unsupported", ""
    mode = mode_of(lit)
    # kernel instantiation: PM_PROJ (4) / PM_AGGM (5) are chosen like jit.cu does
    for i in range(plan.c.n_executors):
        e = plan.c.executors[i]
        if e.tp == ffi.EXEC_PROJECTION and mode == 0:
            mode = 4
        if e.tp in (ffi.EXEC_AGGREGATION, ffi.EXEC_STREAM_AGG) and e.n_group_by > 1:
            mode = 5
    src = ("#define B2_NVRTC 1\n#define B2_JIT_PLAN 1\n#include \"scan_kernel.cuh\"\nnamespace b2 { __constant__ const DevPlan kJitPlan =\n" + lit + ";\n}\n"
           "extern \"C\" __global__ void __launch_bounds__(b2::TILE + 64, 2) b2_scan_jit(const __grid_constant__ b2::ScanArgs A) {\n"
           "  b2::scan_body<" + str(mode) + ">(b2::kJitPlan, A);\n}\n")
    err, prog = nvrtc.nvrtcCreateProgram(src.encode(), b"b2_scan_jit.cu", 0, [], [])
    opts = [b"--gpu-architecture=sm_100a", b"--std=c++17", b"-lineinfo", b"-DB2_NVRTC=1", b"-default-device",
            ("-I" + os.path.join(ROOT, "tikv_b200", "csrc")).encode(), b"-I/usr/local/cuda/include", b"--ptxas-options=-v", b"-DB2_EXT_SIGS=1"]
    (rc,) = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    _, n = nvrtc.nvrtcGetProgramLogSize(prog)
    log = b" " * n
    nvrtc.nvrtcGetProgramLog(prog, log)
    log = log.decode(errors="replace")
    nvrtc.nvrtcDestroyProgram(prog)
    if rc != nvrtc.nvrtcResult.NVRTC_SUCCESS:
        return name, "FAILED", log[:1500]
    facts = " | ".join(l.strip() for l in log.splitlines() if "registers" in l or "spill" in l)
    return name, "ok", facts


def main():
    import scenarios as sc
    plans = []
    for group in ("plans", "int_plans", "minmax_plans", "in_plans", "projection_plans", "multi_group_plans", "scalar_plans"):
        if hasattr(sc, group):
            plans += [(f"{group}:{n}", p) for n, p in getattr(sc, group)()]
    only = sys.argv[1:] 
    bad = 0
    for name, plan in plans:
        if only and not any(o in name for o in only):
            continue
        n, st, facts = compile_plan(plan, name)
        print(f"{st:12s} {n}  {facts}")
        bad += st == "FAILED"
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
