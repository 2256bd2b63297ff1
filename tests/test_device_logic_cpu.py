"""`-m "not gpu"`: the device row logic (b2_device.h) + plan lowering, driven entry-by-entry on the CPU by
tests/host_emul.cpp, must agree with the oracle on every scenario.  The `-m gpu` twin (test_gpu_parity.py) runs the
same scenarios through the CUDA kernels and the C ABI."""
import pytest

import emu
import kvfmt
import orc
import scenarios as sc
from compare import assert_same_rows
from tikv_b200 import ffi
from tikv_b200.plan import ColumnDef, Plan, col, const_int, lt, multiply, plus

PLANS = sc.plans()


@pytest.fixture(scope="module")
def regions():
    return {seed: sc.dirty_region(seed) for seed in (1, 2)}


@pytest.mark.parametrize("name,plan", PLANS, ids=[n for n, _ in PLANS])
@pytest.mark.parametrize("seed", [1, 2])
def test_emulated_device_logic_matches_oracle(name, plan, seed, regions):
    for n_blocks, ranges in ((1, sc.WHOLE), (3, sc.split_ranges())):
        region = regions[seed].build(read_ts=sc.READ_TS, n_write_blocks=n_blocks)
        exp = orc.dag_handle(plan, ranges, region)
        got = emu.dag_handle(plan, ranges, region)
        assert exp.status == 0
        assert_same_rows(got, exp, ordered=not sc.is_agg(name), ctx=f"{name}/seed{seed}/blocks{n_blocks}")
        assert got.stats["processed_keys"] == exp.stats["processed_keys"]
        assert got.stats["processed_size"] == exp.stats["processed_size"]
        assert got.stats["default_lookups"] == exp.stats["data_processed_keys"]
        assert bool(got.stats["met_newer"]) == (exp.stats["met_newer"] == 1)


@pytest.mark.parametrize("name,plan", sc.real_sum_plans())
def test_real_sum_close_to_oracle(name, plan, regions):
    region = regions[1].build(read_ts=sc.READ_TS)
    exp = orc.dag_handle(plan, sc.WHOLE, region)
    got = emu.dag_handle(plan, sc.WHOLE, region)
    assert_same_rows(got, exp, ordered=False, float_rel_tol=1e-12, ctx=name)


def test_real_sums_are_exactly_rounded(regions):
    region = regions[1].build(read_ts=sc.READ_TS)
    sc.check_exact_real_sums(lambda plan: emu.dag_handle(plan, sc.WHOLE, region), region)


def test_isolation_levels_and_read_ts(regions):
    plan = Plan().table_scan(sc.TABLE, sc.COLUMNS).build()
    for ts in (1, 9, 25, 45, 150, sc.READ_TS, sc.READ_TS + 100, (1 << 64) - 1):
        region = regions[2].build(read_ts=ts, isolation=ffi.ISO_RC)
        assert_same_rows(emu.dag_handle(plan, sc.WHOLE, region), orc.dag_handle(plan, sc.WHOLE, region), ctx=f"ts{ts}")
    region = regions[2].build(read_ts=sc.READ_TS, isolation=ffi.ISO_RC_CHECK_TS)
    exp, got = orc.dag_handle(plan, sc.WHOLE, region), emu.dag_handle(plan, sc.WHOLE, region)
    assert exp.status == ffi.B2_ERR_WRITE_CONFLICT == got.status
    assert got.rows() == exp.rows()


def test_checksum_matches_oracle(regions):
    for seed in (1, 2):
        region = regions[seed].build(read_ts=sc.READ_TS, n_write_blocks=2)
        for ranges in (sc.WHOLE, sc.split_ranges()):
            st, exp, _ = orc.checksum(ranges, region)
            st2, got = emu.checksum(ranges, region)
            assert st == 0 == st2 and got == exp and exp[1] > 0
    prefix_old, prefix_new = b"t" + kvfmt.enc_i64_cmp(42), b"t" + kvfmt.enc_i64_cmp(sc.TABLE)
    region = regions[1].build(read_ts=sc.READ_TS)
    st, exp, _ = orc.checksum(sc.WHOLE, region, prefix_old, prefix_new)
    st2, got = emu.checksum(sc.WHOLE, region, prefix_old, prefix_new)
    assert st == 0 == st2 and got == exp
    st, _, msg = orc.checksum(sc.WHOLE, region, b"", b"x")
    st2, _ = emu.checksum(sc.WHOLE, region, b"", b"x")
    assert st != 0 and st2 != 0


def _err_region():
    T = sc.TABLE
    r = kvfmt.Region()
    for h in range(40):
        r.put(kvfmt.row_key(T, h), kvfmt.row_v2([(1, h, "int"), (2, h % 5, "int"), (3, 7, "uint"), (4, 1.0, "f64"), (6, 1, "int")]), 10, 20)
    return r


def test_errors_match_oracle():
    T = sc.TABLE
    cols = sc.COLUMNS
    # corrupted row in the middle: rows before it, then CORRUPTED
    r = _err_region()
    r.put(kvfmt.row_key(T, 20), kvfmt.row_v1([(1, kvfmt.datum_int(5))]) + bytes([kvfmt.VAR_INT, 0x80]), 30, 40)
    plan = Plan().table_scan(T, cols).build()
    region = r.build(read_ts=100)
    exp, got = orc.dag_handle(plan, sc.WHOLE, region), emu.dag_handle(plan, sc.WHOLE, region)
    assert exp.status == ffi.B2_ERR_CORRUPTED == got.status and got.rows() == exp.rows() and len(exp.rows()) == 20
    # v2 int with a 3-byte width
    r = _err_region()
    bad = bytearray(kvfmt.row_v2([(1, 70000, "int"), (2, 1, "int")]))
    bad[8] = 3  # first end-offset: value 1 becomes 3 bytes wide
    r.put(kvfmt.row_key(T, 7), bytes(bad), 30, 40)
    region = r.build(read_ts=100)
    exp, got = orc.dag_handle(plan, sc.WHOLE, region), emu.dag_handle(plan, sc.WHOLE, region)
    assert exp.status == ffi.B2_ERR_CORRUPTED == got.status and got.rows() == exp.rows()
    # BIGINT overflow in an expression -> Evaluate error 1690
    r = _err_region()
    r.put(kvfmt.row_key(T, 1000), kvfmt.row_v2([(1, (1 << 62), "int"), (2, 4, "int"), (3, 7, "uint"), (6, 1, "int")]), 10, 20)
    p2 = Plan().table_scan(T, cols).selection(lt(multiply(col(sc.C1), const_int(4)), const_int(100))).build()
    region = r.build(read_ts=100)
    exp, got = orc.dag_handle(p2, sc.WHOLE, region), emu.dag_handle(p2, sc.WHOLE, region)
    assert exp.status == ffi.B2_ERR_EVALUATE == got.status and exp.mysql_code == 1690
    # bad write record / missing default CF value -> Storage error
    r = _err_region().raw_write(kvfmt.row_key(T, 5), 50, b"Xjunk")
    region = r.build(read_ts=100)
    exp, got = orc.dag_handle(plan, sc.WHOLE, region), emu.dag_handle(plan, sc.WHOLE, region)
    assert exp.status == ffi.B2_ERR_STORAGE == got.status and got.rows() == exp.rows()
    r = _err_region()
    r.write.append((kvfmt.write_key(kvfmt.row_key(T, 3), 60), kvfmt.write_record(b"P", 55)))  # long value without CF_DEFAULT entry
    region = r.build(read_ts=100)
    exp, got = orc.dag_handle(plan, sc.WHOLE, region), emu.dag_handle(plan, sc.WHOLE, region)
    assert exp.status == ffi.B2_ERR_STORAGE == got.status and got.rows() == exp.rows()
    # decode error only matters for rows that are still selected (lazy decode)
    r = _err_region()
    r.put(kvfmt.row_key(T, 500), kvfmt.row_v1([(1, kvfmt.datum_int(1)), (2, kvfmt.datum_bytes(b"zz")), (6, kvfmt.datum_int(1))]), 10, 20)
    p3 = Plan().table_scan(T, cols).selection(lt(col(sc.C_H), const_int(100))).build(output_offsets=[sc.C_H, sc.C2])
    region = r.build(read_ts=100)
    exp, got = orc.dag_handle(p3, sc.WHOLE, region), emu.dag_handle(p3, sc.WHOLE, region)
    assert exp.status == 0 == got.status and got.rows() == exp.rows()
    p4 = Plan().table_scan(T, cols).build(output_offsets=[sc.C_H, sc.C2])
    exp, got = orc.dag_handle(p4, sc.WHOLE, region), emu.dag_handle(p4, sc.WHOLE, region)
    assert exp.status == ffi.B2_ERR_CORRUPTED == got.status


def test_check_supported_mirrors_runner():
    cols = [ColumnDef(1, pk_handle=True), ColumnDef(2), ColumnDef(3, tp=ffi.TP_VARCHAR)]
    ok = Plan().table_scan(5, cols).selection(lt(col(1), const_int(3))).build(output_offsets=[0, 1])
    assert emu.check_supported(ok)[0] == 0
    assert emu.check_supported(Plan().table_scan(5, cols).build())[0] == ffi.B2_OK  # varchar output: materialised since ABI 3
    assert emu.check_supported(Plan().table_scan(5, cols + [ColumnDef(4, tp=ffi.TP_ENUM)]).build())[0] == ffi.B2_ERR_UNSUPPORTED  # enum output
    assert emu.check_supported(Plan().table_scan(5, cols, desc=True).build(output_offsets=[0]))[0] == ffi.B2_OK  # backward scans are on the device path
    two = Plan().table_scan(5, cols).aggregation([("count", const_int(1))], group_by=[col(0), col(1)]).build()
    assert emu.check_supported(two)[0] == 0  # BatchSlowHashAggregation: up to 4 Int / Real expressions
    five = Plan().table_scan(5, cols).aggregation([("count", const_int(1))], group_by=[col(0), col(1), col(0), col(1), col(0)]).build()
    assert emu.check_supported(five)[0] == ffi.B2_ERR_UNSUPPORTED
    rc, msg = emu.check_supported(Plan().table_scan(5, cols).selection(lt(col(2), const_int(3))).build(output_offsets=[0]))
    assert rc == ffi.B2_ERR_UNSUPPORTED and "eval type does not match" in msg


@pytest.mark.parametrize("name,plan,exact,keys", sc.topn_plans(), ids=[t[0] for t in sc.topn_plans()])
def test_topn_device_logic_matches_oracle(name, plan, exact, keys, regions):
    from compare import assert_topn
    for seed, n_blocks, ranges in ((1, 1, sc.WHOLE), (2, 3, sc.split_ranges())):
        region = regions[seed].build(read_ts=sc.READ_TS, n_write_blocks=n_blocks)
        assert_topn(emu.dag_handle(plan, ranges, region), orc.dag_handle(plan, ranges, region), exact, keys, ctx=f"{name}/seed{seed}")


@pytest.mark.parametrize("fmt", [2, 1])
@pytest.mark.parametrize("name,plan", sc.int_plans(), ids=[n for n, _ in sc.int_plans()])
def test_exact_layout_fast_path(name, plan, fmt):
    """All-integer table, row formats v2 and v1: layout probes, `column <cmp> const` conditions and outputs by stored
    position, every width / datum-flag mix, signed/unsigned compares; rows with NULL / missing / extra / reordered
    columns take the general path in the same batch."""
    region = sc.int_region(3, fmt=fmt).build(read_ts=sc.READ_TS, n_write_blocks=2)
    exp = orc.dag_handle(plan, sc.WHOLE, region)
    got = emu.dag_handle(plan, sc.WHOLE, region)
    assert exp.status == 0 and exp.n_rows > 0
    if name == "topn":
        from compare import assert_topn
        assert_topn(got, exp, True, None, ctx=name)
    else:
        assert_same_rows(got, exp, ordered=name != "agg", ctx=name)


@pytest.mark.parametrize("ts", [(5, 8), (449_000_000_000_000_001, 449_000_000_000_000_777), ((1 << 63) + 5, (1 << 63) + 9)],
                         ids=["ts1byte", "ts9bytes", "ts10bytes"])
@pytest.mark.parametrize("name,plan", sc.int_plans(), ids=[n for n, _ in sc.int_plans()])
def test_clean_entry_front_end(name, plan, ts):
    """entry_fast (word-wise key tail / write head / v2 row) accepts the clean entries and yields what the general walk
    yields: same result with it switched off, same as the oracle.  start_ts as TiDB issues it is a 9-byte varint; a
    10-byte one is left to the general parser."""
    region = sc.int_region(5, ts=ts).build(read_ts=ts[1] + 10, n_write_blocks=2)
    exp = orc.dag_handle(plan, sc.WHOLE, region)
    before = emu.fast_hits()
    got = emu.dag_handle(plan, sc.WHOLE, region)
    hits = emu.fast_hits() - before
    emu.set_fast_front(False)
    try:
        slow = emu.dag_handle(plan, sc.WHOLE, region)
    finally:
        emu.set_fast_front(True)
    assert exp.status == 0 and exp.n_rows > 0
    assert hits > (400 if ts[0] < (1 << 63) else -1), hits  # ~92 % of the 500 rows have the exact layout
    if ts[0] >= (1 << 63):
        assert hits == 0
    assert got.stats == slow.stats
    if name == "topn":
        from compare import assert_topn
        assert_topn(got, exp, True, None, ctx=name)
        assert_topn(slow, exp, True, None, ctx=name)
    else:
        assert_same_rows(got, exp, ordered=name != "agg", ctx=name)
        assert_same_rows(slow, exp, ordered=name != "agg", ctx=name)


@pytest.mark.parametrize("fmt", [2, 1])
def test_exact_layout_fast_path_corrupted_rows(fmt):
    """v2: 3/5/9-byte integers and decreasing offsets; v1: truncated datums and dangling column markers.  The probe must
    reject the row and the general path must raise the reference's error at the same entry with the rows before it intact."""
    region = sc.int_region(4, corrupt=True, fmt=fmt).build(read_ts=sc.READ_TS)
    plan = sc.int_plans()[1][1]
    exp = orc.dag_handle(plan, sc.WHOLE, region)
    got = emu.dag_handle(plan, sc.WHOLE, region)
    assert exp.status != 0 and got.status == exp.status
    assert_same_rows(got, exp, ordered=True, ctx="corrupt")


@pytest.mark.parametrize("name,plan", sc.limit_plans(), ids=[n for n, _ in sc.limit_plans()])
def test_limit(name, plan, regions):
    region = regions[1].build(read_ts=sc.READ_TS, n_write_blocks=2)
    exp = orc.dag_handle(plan, sc.split_ranges(), region)
    got = emu.dag_handle(plan, sc.split_ranges(), region)
    assert_same_rows(got, exp, ordered=True, ctx=name)


def test_wordwise_varints_match_oracle():
    """dec_var_u64 / first_var_int_len (one 8-byte load + bit tricks) against the oracle's byte loop: every length 1..10,
    truncated buffers, over-long encodings (codec number.rs:445-567)."""
    import ctypes as C
    import random
    E, O = emu.lib(), orc.lib()
    E.emu_dec_var_u64.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_uint64)]
    E.emu_first_var_int_len.argtypes = [C.c_char_p, C.c_uint32]
    E.emu_split_datum.argtypes = [C.c_char_p, C.c_uint32]
    rng = random.Random(11)
    cases = []
    for bits in list(range(0, 65)) * 4:
        v = rng.getrandbits(bits) if bits else 0
        cases.append(kvfmt.enc_var_u64(v))
    cases += [bytes([0x80] * k + [0x01]) for k in range(0, 12)] + [bytes([0xFF] * k) for k in range(1, 13)] + [bytes([0xFF] * 9 + [b]) for b in (0, 1, 2, 0x7F, 0x80, 0xFF)]
    for enc in cases:
        for n in range(0, len(enc) + 1):
            buf = enc[:n] + bytes(rng.getrandbits(8) for _ in range(16))  # garbage after the slice must not matter
            a, b = C.c_uint64(0), C.c_uint64(0)
            ra = E.emu_dec_var_u64(buf, n, C.byref(a))
            rb = O.orc_decode_var_u64(buf, n, C.byref(b))
            assert ra == rb and (ra == 0 or a.value == b.value), (enc.hex(), n, ra, rb, a.value, b.value)
            datum = bytes([8]) + buf
            assert E.emu_split_datum(datum, n + 1) == O.orc_split_datum(datum, n + 1), (enc.hex(), n)


@pytest.mark.parametrize("name,plan", sc.minmax_plans(), ids=[n for n, _ in sc.minmax_plans()])
def test_min_max(name, plan, regions):
    region = regions[2].build(read_ts=sc.READ_TS, n_write_blocks=2)
    exp = orc.dag_handle(plan, sc.WHOLE, region)
    got = emu.dag_handle(plan, sc.WHOLE, region)
    assert exp.status == 0
    assert_same_rows(got, exp, ordered=False, ctx=name)


@pytest.mark.parametrize("name,plan", sc.multi_group_plans(), ids=[n for n, _ in sc.multi_group_plans()])
def test_multi_column_group_by(name, plan, regions):
    for seed in (1, 2):
        region = regions[seed].build(read_ts=sc.READ_TS, n_write_blocks=2)
        exp = orc.dag_handle(plan, sc.WHOLE, region)
        got = emu.dag_handle(plan, sc.WHOLE, region)
        assert exp.status == 0 and (exp.n_rows > 0 or name == "mg_no_input")
        assert_same_rows(got, exp, ordered=False, float_rel_tol=1e-12 if name == "mg_same_expr_twice" else None, ctx=f"{name}/seed{seed}")


@pytest.mark.parametrize("name,plan", sc.scalar_plans(), ids=[n for n, _ in sc.scalar_plans()])
def test_scalar_functions(name, plan, regions):
    for seed in (1, 2):
        region = regions[seed].build(read_ts=sc.READ_TS, n_write_blocks=2)
        exp = orc.dag_handle(plan, sc.WHOLE, region)
        got = emu.dag_handle(plan, sc.WHOLE, region)
        if "_err_" in name:  # evaluation errors end the request (the oracle's rows are a prefix, see test_projection)
            assert exp.status == ffi.B2_ERR_EVALUATE == got.status and exp.mysql_code == 1690, (name, exp.status, exp.message)
            assert got.rows()[:len(exp.rows())] == exp.rows()
            continue
        assert exp.status == 0 and exp.n_rows > 0, exp.message
        assert_same_rows(got, exp, ordered=not sc.is_agg(name) and "agg" not in name, ctx=f"{name}/seed{seed}")


@pytest.mark.parametrize("name,plan", sc.in_plans(), ids=[n for n, _ in sc.in_plans()])
def test_in_lists(name, plan, regions):
    region = regions[1].build(read_ts=sc.READ_TS, n_write_blocks=2)
    exp = orc.dag_handle(plan, sc.WHOLE, region)
    got = emu.dag_handle(plan, sc.WHOLE, region)
    assert exp.status == 0 and exp.n_rows > 0
    assert_same_rows(got, exp, ordered=not sc.is_agg(name) and "group" not in name, ctx=name)


@pytest.mark.parametrize("name,plan", sc.projection_plans(), ids=[n for n, _ in sc.projection_plans()])
def test_projection(name, plan, regions):
    region = sc.dirty_region(1, full_range=name == "proj_overflow").build(read_ts=sc.READ_TS, n_write_blocks=2)
    exp = orc.dag_handle(plan, sc.WHOLE, region)
    got = emu.dag_handle(plan, sc.WHOLE, region)
    if name == "proj_overflow":
        # an evaluation error ends the request; the reference drops the rows of the batch it happened in (projection_executor.rs
        # :207-211, batch = 32..1024 rows), the device keeps every row before the failing one: the oracle's rows are a prefix
        assert exp.status == ffi.B2_ERR_EVALUATE == got.status and exp.mysql_code == 1690
        assert got.rows()[:len(exp.rows())] == exp.rows()
        return
    assert exp.status == 0
    assert_same_rows(got, exp, ordered=True, ctx=name)


def test_scalar_function_known_answers():
    sc.check_scalar_known_answers(emu.dag_handle)


def test_like_known_answers():
    """impl_like.rs test_like / test_like_wide_character through the device logic (like_match, plan lowering, the bytes
    constants' pool); _ci collations are reported unsupported."""
    sc.check_like_known_answers(emu.dag_handle)
    from tikv_b200.plan import ColumnDef, Plan, const_bytes, like
    cols = [ColumnDef(100, pk_handle=True), ColumnDef(1)]
    ci = Plan().table_scan(sc.TABLE, cols).projection(like(const_bytes("ßssß".encode(), -45), const_bytes("_sSß".encode(), -45), collation=-45)).build()
    rc, msg = emu.check_supported(ci)
    assert rc == ffi.B2_ERR_UNSUPPORTED and "collation" in msg


def test_plan_limits_are_reported_not_crashed():
    """Plans beyond the device path's static limits (plan_compile.h) come back as B2_ERR_UNSUPPORTED / INVALID_ARG with a
    message — the host then keeps the CPU executors — instead of overrunning a fixed-size table."""
    from tikv_b200.plan import Expr, case_when, fn, if_, plus
    scan = lambda: Plan().table_scan(sc.TABLE, sc.COLUMNS)
    c2, c6 = col(sc.C2), col(sc.C6, tp=ffi.TP_LONG)
    deep = c2
    for _ in range(40):  # left-deep: few stack slots, many nodes
        deep = plus(deep, const_int(1))
    right = const_int(0)
    for _ in range(20):  # right-deep: the RPN stack grows with every level
        right = plus(c2, right)
    cases = {
        "too many nodes": scan().selection(*[lt(deep, const_int(5)) for _ in range(3)]).build(),
        "expression too deep": scan().selection(lt(right, const_int(5))).build(),
        "nine conditions": scan().selection(*[lt(c2, const_int(i)) for i in range(9)]).build(),
        "nine aggregates": scan().aggregation([("count", c2)] * 9).build(),
        "five group-by": scan().aggregation([("count", c2)], group_by=[c2, c6, c2, c6, c2]).build(),
        "five order-by": scan().topn([(c2, False)] * 5, 10).build(),
        "topn limit": scan().topn([(c2, False)], 5000).build(),
        "17 projections": scan().projection(*[c2] * 17).build(),
        "if arity": scan().selection(fn("IF_INT", c2, c6)).build(),
        "case_when cond type": scan().selection(case_when(col(sc.C4, tp=ffi.TP_DOUBLE), c2)).build(),
        "cmp arg type": scan().selection(fn("LT_INT", c2, col(sc.C4, tp=ffi.TP_DOUBLE))).build(),
        "compare arity": scan().selection(fn("LT_INT", c2)).build(),
        "unknown sig": scan().selection(Expr(ffi.RPN_FN, ffi.TP_LONGLONG, sig=999999, args=(c2,))).build(),
        "varchar NULL constant": scan().selection(lt(c2, Expr(ffi.RPN_CONST_NULL, ffi.TP_VARCHAR))).build(),
        "output offset": scan().build(output_offsets=[99]),
        "projection then selection": scan().projection(c2).selection(lt(col(0), const_int(1))).build(),
    }
    for what, plan in cases.items():
        rc, msg = emu.check_supported(plan)
        assert rc in (ffi.B2_ERR_UNSUPPORTED, ffi.B2_ERR_INVALID_ARG) and msg, (what, rc, msg)
    # and the limits themselves are usable
    ok = {
        "eight conditions": scan().selection(*[lt(c2, const_int(i)) for i in range(8)]).build(),
        "four group-by": scan().aggregation([("count", c2)], group_by=[c2, c6, c2, c6]).build(),
        "16 projections": scan().projection(*[c2] * 16).build(),
        "if": scan().selection(if_(c2, c6, c2)).build(),
    }
    for what, plan in ok.items():
        rc, msg = emu.check_supported(plan)
        assert rc == 0, (what, msg)


@pytest.mark.parametrize("seed", range(12))
def test_run_ownership_of_the_lean_front_end(seed):
    """Version runs of every length (1..90 entries: inside a warp, across one or several 32-entry warps), made of newer
    versions, Puts, Deletes, Lock / Rollback records and long values, read at a random snapshot under SI / RC / RcCheckTs:
    the lean front end (32 lanes at a time, b2_device.h fast_lane_decide) commits the plain runs and pushes every other
    run start exactly once; the result equals the oracle's and the general walk's."""
    import random
    rng = random.Random(1000 + seed)
    r = kvfmt.Region()
    mk = lambda: kvfmt.row_v2([(cid, rng.randrange(-(1 << 40), 1 << 40) if not uns else rng.randrange(0, 1 << 40), "uint" if uns else "int")
                               for cid, uns in ((2, False), (3, True), (4, False), (5, False), (6, True), (7, True), (9, False), (11, False))])
    for h in range(160):
        key = kvfmt.row_key(sc.TABLE, h * 2 + 5)
        n = rng.choice([1, 1, 1, 2, 3, 5, 9, 31, 32, 33, 40, 70, 90]) if rng.random() < 0.5 else 1
        ts = 1000
        for _ in range(n):
            ts -= rng.randrange(2, 9)
            x = rng.random()
            if x < 0.62:
                r.put(key, mk(), ts - 1, ts)
            elif x < 0.72:
                r.delete(key, ts - 1, ts)
            elif x < 0.84:
                r.lock_rec(key, ts - 1, ts) if rng.random() < 0.5 else r.lock_rec(key, ts - 1, ts, last_change=(ts - rng.randrange(3, 40), rng.randrange(1, 12)))
            elif x < 0.90:
                r.rollback(key, ts)
            else:
                r.put(key, mk(), ts - 1, ts, force_long=True)
    plans = dict(sc.int_plans())
    for iso in (ffi.ISO_SI, ffi.ISO_RC, ffi.ISO_RC_CHECK_TS):
        for read_ts in (rng.randrange(300, 1000), 1001, 5):
            region = r.build(read_ts=read_ts, n_write_blocks=rng.choice([1, 2, 3]), isolation=iso)
            for name in ("all", "agg"):
                exp = orc.dag_handle(plans[name], sc.WHOLE, region)
                before = emu.fast_hits()
                got = emu.dag_handle(plans[name], sc.WHOLE, region)
                assert got.status == exp.status, (name, iso, read_ts, got.message)
                assert read_ts != 1001 or exp.status != 0 or emu.fast_hits() - before > 20  # most runs are plain at the newest snapshot
                assert_same_rows(got, exp, ordered=name != "agg", ctx=f"{name}/iso{iso}/ts{read_ts}")
                assert got.stats["processed_keys"] == exp.stats["processed_keys"] and got.stats["processed_size"] == exp.stats["processed_size"]
                assert bool(got.stats["met_newer"]) == (exp.stats["met_newer"] == 1)
            st, exp_ck, _ = orc.checksum(sc.WHOLE, region)
            st2, got_ck = emu.checksum(sc.WHOLE, region)
            assert (st == 0) == (st2 == 0)
            if st == 0:
                assert got_ck == exp_ck
