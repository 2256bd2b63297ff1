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
