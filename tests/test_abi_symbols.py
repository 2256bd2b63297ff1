"""`-m "not gpu"`: the C-ABI library builds for sm_100a, loads, and exports every symbol include/b2_copr.h declares.
No compute call is made (there is no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from tikv_b200 import ffi
    return ffi.lib()


def declared_functions():
    src = open(os.path.join(ROOT, "include", "b2_copr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in b2_copr.h but not exported by libb2copr.so"
    from tikv_b200 import ffi
    assert set(ffi.EXPORTED_SYMBOLS) == set(names)


def test_abi_version_and_structs(lib):
    from tikv_b200 import ffi
    assert lib.b2_abi_version() == 4
    assert b"sm_100a" in lib.b2_build_info()
    # struct sizes the Rust/cgo side would mirror
    assert C.sizeof(ffi.Decimal) == 40 and C.sizeof(ffi.CfBlock) == 40 and C.sizeof(ffi.RpnNode) == 40
    assert C.sizeof(ffi.Column) == 48 and C.sizeof(ffi.ColumnInfo) == 40 and C.sizeof(ffi.ChecksumResponse) == 24


def test_sass_is_sm100a(lib):
    """The shipped library must contain sm_100a SASS for the hot kernels (no PTX-only / other-arch fallback)."""
    import subprocess
    so = os.path.join(ROOT, "tikv_b200", "_build", "libb2copr.so")
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_check_supported_without_gpu(lib):
    from tikv_b200 import ffi
    from tikv_b200.plan import ColumnDef, Plan, col, const_int, lt
    cols = [ColumnDef(1, pk_handle=True), ColumnDef(2), ColumnDef(3, tp=ffi.TP_VARCHAR)]
    ok = Plan().table_scan(5, cols).selection(lt(col(1), const_int(3))).build(output_offsets=[0, 1])
    assert lib.b2_check_supported(C.byref(ok.c)) == ffi.B2_OK
    assert lib.b2_check_supported(C.byref(Plan().table_scan(5, cols).build().c)) == ffi.B2_OK  # VARCHAR output: materialised since ABI 3
    bad = Plan().table_scan(5, cols).selection(lt(col(2), const_int(3))).build()
    assert lib.b2_check_supported(C.byref(bad.c)) == ffi.B2_ERR_UNSUPPORTED
    assert b"eval type does not match" in lib.b2_last_error_message()  # (a VARCHAR operand is a LIKE operand only)
    # DATETIME / DURATION columns take part in comparisons (ABI 4); a time-valued expression is still the CPU's
    from tikv_b200.plan import const_time
    tcols = [ColumnDef(1, pk_handle=True), ColumnDef(2, tp=ffi.TP_DATETIME)]
    assert lib.b2_check_supported(C.byref(Plan().table_scan(5, tcols).selection(lt(col(1, tp=ffi.TP_DATETIME), const_time(1 << 40))).build().c)) == ffi.B2_OK
    assert lib.b2_check_supported(C.byref(Plan().table_scan(5, tcols).selection(col(1, tp=ffi.TP_DATETIME)).build().c)) == ffi.B2_ERR_UNSUPPORTED


def test_open_fails_loudly_without_cuda(lib):
    """No CPU fallback: without a CUDA device b2_exec_open reports B2_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import kvfmt
    from tikv_b200 import ffi
    from tikv_b200.executor import B2Error, BatchExecutor
    from tikv_b200.plan import ColumnDef, Plan
    r = kvfmt.Region().put(kvfmt.row_key(5, 1), kvfmt.row_v2([(2, 1, "int")]), 1, 2).build(read_ts=10)
    p = Plan().table_scan(5, [ColumnDef(1, pk_handle=True), ColumnDef(2)]).build()
    with pytest.raises(B2Error) as ei:
        BatchExecutor(p, [kvfmt.table_range(5)], r)
    assert ei.value.status == ffi.B2_ERR_CUDA and "no CPU fallback" in ei.value.message


def test_header_is_plain_c(tmp_path):
    """include/b2_copr.h is what cgo / bindgen / a C host would consume: it must compile as C11 with -pedantic."""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "b2_copr.h"\nint main(void) { return (int)sizeof(b2_exec_config) == 0; }\n')
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_ctypes_mirror_matches_header_layout(tmp_path):
    """Every struct of include/b2_copr.h against its ctypes mirror in tikv_b200/ffi.py: total size and the offset of every
    field, taken from a C program compiled against the header (what bindgen would see)."""
    import subprocess
    from tikv_b200 import ffi
    pairs = {"b2_cf_block": ffi.CfBlock, "b2_region_source": ffi.RegionSource, "b2_key_range": ffi.KeyRange, "b2_column_info": ffi.ColumnInfo,
             "b2_rpn_node": ffi.RpnNode, "b2_rpn_expr": ffi.RpnExpr, "b2_aggr_desc": ffi.AggrDesc, "b2_order_by": ffi.OrderBy,
             "b2_executor_desc": ffi.ExecutorDesc, "b2_dag_plan": ffi.DagPlan, "b2_exec_config": ffi.ExecConfig, "b2_decimal": ffi.Decimal,
             "b2_column": ffi.Column, "b2_batch": ffi.Batch, "b2_exec_stats": ffi.ExecStats, "b2_error_info": ffi.ErrorInfo,
             "b2_checksum_response": ffi.ChecksumResponse, "b2_agg_partials": ffi.AggPartials, "b2_gen_spec": ffi.GenSpec,
             "b2_gen_block": ffi.GenBlock, "b2_encoded_chunk": ffi.EncodedChunk, "b2_sst_blocks": ffi.SstBlocks, "b2_sst_stats": ffi.SstStats,
             "b2_sst_encoded": ffi.SstEncoded}
    header = open(os.path.join(ROOT, "include", "b2_copr.h")).read()
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "b2_copr.h"', 'int main(void) {']
    expect = {}
    for cname, mirror in pairs.items():
        m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), header, re.S)
        assert m, cname
        body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
        names = [re.sub(r"\[.*\]", "", part.strip().split()[-1].lstrip("*")) for d in body.split(";") if d.strip() for part in d.split(",")]
        assert names == [f[0] for f in mirror._fields_], (cname, names, [f[0] for f in mirror._fields_])
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for n in names:
            lines.append(f'  printf(" %zu", offsetof({cname}, {n}));')
        lines.append('  printf("\\n");')
        expect[cname] = [C.sizeof(mirror)] + [getattr(mirror, n).offset for n in names]
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True)
    got = {l.split()[0]: [int(x) for x in l.split()[1:]] for l in out.splitlines()}
    assert got == expect


def test_integration_doc_structs_follow_header():
    """The #[repr(C)] structs shown in INTEGRATION.md list the header's fields in the header's order (a trailing `_` is
    allowed where the C name is a Rust type name), and every entry point of the header is bound or named there."""
    header = open(os.path.join(ROOT, "include", "b2_copr.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    n = 0
    for m in re.finditer(r"pub struct (b2_\w+) \{(.*?)\}", doc, re.S):
        name, body = m.group(1), m.group(2)
        fields = [re.sub(r"^pub\s+", "", f.strip()).split(":")[0].strip().rstrip("_") or "_" for f in re.split(r",(?![^\[]*\])", body) if ":" in f]
        h = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), header, re.S)
        assert h, name
        hb = re.sub(r"/\*.*?\*/", "", h.group(1), flags=re.S)
        hf = [re.sub(r"\[.*\]", "", part.strip().split()[-1].lstrip("*")) for d in hb.split(";") if d.strip() for part in d.split(",")]
        assert [f if f.startswith("_") else f for f in fields] == [x.rstrip("_") if not x.startswith("_") else x for x in hf], (name, fields, hf)
        n += 1
    assert n >= 12
    for fn_name in set(re.findall(r"\b(b2_[a-z_]+)\(", header)):
        assert fn_name in doc, fn_name


def test_product_path_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under tikv_b200/ (the shipped package and its CUDA sources) may import,
    link or mention it, except smoke.py, which is __graft_entry__.smoke()'s checker; libb2copr.so must not depend on it."""
    import subprocess
    pat = re.compile(r"\boracle\b|liborc|\bimport orc\b|\borc\.")
    pkg = os.path.join(ROOT, "tikv_b200")
    offenders = []
    for d, _, fs in os.walk(pkg):
        if "_build" in d or "__pycache__" in d:
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")) and f != "smoke.py":
                text = open(os.path.join(d, f), errors="replace").read()
                if pat.search(text):
                    offenders.append(os.path.relpath(os.path.join(d, f), ROOT))
    assert not offenders, offenders
    so = os.path.join(pkg, "_build", "libb2copr.so")
    needed = subprocess.check_output(["readelf", "-d", so], text=True)
    assert "liborc" not in needed and "libemu" not in needed


def test_committed_bench_lines_follow_the_contract():
    """profiles/bench_r1.json and bench_reference_r1.json (the lines bench.py printed on the B200) carry every key of the
    bench contract with sane types; the roofline fraction is achieved / peak."""
    import json
    ours = json.load(open(os.path.join(ROOT, "profiles", "bench_r1.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "roofline", "e2e", "cpu_baseline", "gpu_launches", "clocks"):
        assert k in ours, k
    assert ours["vs_baseline"] is None and ours["higher_is_better"] is True and ours["scaling"] == "weak" and "workload" in ours["config"]
    rf = ours["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and rf["traffic"]
    e = ours["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < ours["value"]
    cb = ours["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] > 0
    assert ours["gpu_launches"] > 0 and not ours["clocks"]["reasons"] and ours["clocks"]["sm_mhz"] > 0
    ref = json.load(open(os.path.join(ROOT, "profiles", "bench_reference_r1.json")))
    assert ref["impl"] == "reference" and ref["metric"] == ours["metric"] and ref["unit"] == ours["unit"] and ref["config"]["workload"].startswith("C2: BatchTableScan + BatchSelection(col0 < 0)")
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 == ref["e2e"]["d2h_bytes_per_step"] and ref["e2e"]["value"] == ref["value"] == ref["cpu_baseline"]["value"]


def test_scalar_function_numbers_follow_tipb():
    """b2_rpn_node.sig carries tipb::ScalarFuncSig unchanged.  tipb (pingcap/tipb @ 1374320b, Cargo.lock) is not vendored in
    the reference tree, so the numbers of expression.proto are restated here once, by name, and every B2_SIG_* of the
    header must agree with them (five values had drifted)."""
    from tikv_b200 import ffi
    tipb = {"CastIntAsInt": 0, "CastIntAsReal": 1, "CastRealAsReal": 11,
            "LTInt": 100, "LTReal": 101, "LTDecimal": 102, "LEDecimal": 112, "GTDecimal": 122, "GEDecimal": 132, "EQDecimal": 142, "NEDecimal": 152, "NullEQDecimal": 162,
            "DecimalIsNull": 3111, "InDecimal": 4003, "LTTime": 104, "LTDuration": 105, "LEInt": 110, "LEReal": 111, "LETime": 114, "LEDuration": 115,
            "GTInt": 120, "GTReal": 121, "GTTime": 124, "GTDuration": 125, "GEInt": 130, "GEReal": 131, "GETime": 134, "GEDuration": 135,
            "EQInt": 140, "EQReal": 141, "EQTime": 144, "EQDuration": 145, "NEInt": 150, "NEReal": 151, "NETime": 154, "NEDuration": 155,
            "NullEQInt": 160, "NullEQReal": 161, "NullEQTime": 164, "NullEQDuration": 165,
            "PlusReal": 200, "PlusInt": 203, "MinusReal": 204, "MinusInt": 207, "MultiplyReal": 208, "MultiplyInt": 210, "DivideReal": 211,
            "IntDivideInt": 213, "ModReal": 215, "ModInt": 217, "MultiplyIntUnsigned": 218,
            "AbsInt": 2101, "AbsUInt": 2102, "AbsReal": 2103,
            "LogicalAnd": 3101, "LogicalOr": 3102, "LogicalXor": 3103, "UnaryNotInt": 3104, "UnaryNotReal": 3106, "UnaryMinusInt": 3108, "UnaryMinusReal": 3109,
            "DurationIsNull": 3112, "RealIsNull": 3113, "TimeIsNull": 3115, "IntIsNull": 3116, "BitAndSig": 3118, "BitOrSig": 3119, "BitXorSig": 3120, "BitNegSig": 3121,
            "IntIsTrue": 3122, "RealIsTrue": 3123, "IntIsFalse": 3125, "RealIsFalse": 3126,
            "InInt": 4001, "InReal": 4002, "InTime": 4005, "InDuration": 4006,
            "IfNullInt": 4101, "IfNullReal": 4102, "IfInt": 4107, "IfReal": 4108, "CoalesceInt": 4201, "CoalesceReal": 4202, "CaseWhenInt": 4208, "CaseWhenReal": 4209, "LikeSig": 4310}

    def b2_name(n):
        n = {"BitAndSig": "BitAnd", "BitOrSig": "BitOr", "BitXorSig": "BitXor", "BitNegSig": "BitNeg", "AbsUInt": "AbsUint", "LikeSig": "Like"}.get(n, n)
        for a, b in (("LT", "Lt"), ("LE", "Le"), ("GT", "Gt"), ("GE", "Ge"), ("EQ", "Eq"), ("NE", "Ne")):
            if n.startswith(a) and n[2:3].isupper():
                n = b + n[2:]
        n = n.replace("NullEQ", "Nulleq")
        return re.sub(r"(?<!^)(?=[A-Z])", "_", n).upper()
    names = {b2_name(k): v for k, v in tipb.items()}
    assert set(names) == set(ffi.SIG), (sorted(set(names) ^ set(ffi.SIG)))
    for k, v in names.items():
        assert ffi.SIG[k] == v, (k, ffi.SIG[k], v)
