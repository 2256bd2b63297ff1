"""The plan-specialised kernels must compile under NVRTC (no <stdint.h> macros, no host headers): compile two small
representative plans here, without a GPU, exactly as jit.cu does at run time (tools/jit_compile_check.py compiles every
scenario plan; a composite-key aggregation alone takes over a minute on this box, so it stays out of the test suite)."""
import os
import sys

import pytest

import scenarios as sc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_specialised_kernels_compile_with_nvrtc():
    pytest.importorskip("cuda.bindings.nvrtc")
    import jit_compile_check as J
    from tikv_b200.plan import Plan, col, const_int, gt, if_null, mod
    ext = Plan().table_scan(sc.TABLE, sc.COLUMNS).selection(gt(mod(if_null(col(sc.C2), const_int(3)), const_int(7)), const_int(2))).build(output_offsets=[sc.C_H])
    cases = [("sel_lt_const", dict(sc.plans())["sel_lt_const"]),  # scan + selection (the bench shape)
             ("ext_selection", ext)]                               # the extended scalar functions (B2_EXT_SIGS=1)
    for name, plan in cases:
        n, status, facts = J.compile_plan(plan, name)
        assert status == "ok", (name, facts)
        assert "registers" in facts


def test_literals_are_launch_parameters_and_kernels_are_cached_on_disk():
    """`col < 0` and `col < 1` (and another TopN LIMIT, another IN list) are one plan shape: the device plan keeps slot
    numbers, the values travel in ScanArgs::imms.  The second shape-equal plan compiles nothing: its kernel is found in
    the on-disk cache (b2_plan_precompile needs NVRTC only, no GPU)."""
    import ctypes as C
    pytest.importorskip("cuda.bindings.nvrtc")
    import jit_compile_check as J
    from tikv_b200 import ffi
    from tikv_b200.plan import Plan, col, const_int, in_, lt
    mk = lambda k: Plan().table_scan(sc.TABLE, sc.COLUMNS).selection(lt(col(sc.C2), const_int(k))).build(output_offsets=[sc.C_H])
    assert J.literal_of(mk(0)) == J.literal_of(mk(1)) == J.literal_of(mk(-(1 << 63)))
    topn = lambda n: Plan().table_scan(sc.TABLE, sc.COLUMNS).topn([(col(sc.C2), True)], n).build()
    assert J.literal_of(topn(10)) == J.literal_of(topn(1000))
    inl = lambda *v: Plan().table_scan(sc.TABLE, sc.COLUMNS).selection(in_(col(sc.C2), *[const_int(x) for x in v])).build()
    assert J.literal_of(inl(1, 2, 3)) == J.literal_of(inl(7, 8, 9)) != J.literal_of(inl(1, 2))
    L = ffi.lib()
    n = C.c_int32(-1)
    assert L.b2_plan_precompile(C.byref(mk(5).c), C.byref(n)) == 0, L.b2_last_error_message()
    assert L.b2_plan_precompile(C.byref(mk(6).c), C.byref(n)) == 0
    assert n.value == 0  # same shape: served by the cache
