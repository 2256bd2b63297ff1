"""`-m "not gpu"`: LIKE and DECIMAL comparisons of the oracle and of the device logic (compiled for the host) against two
independent references: Python's `re` (a LIKE pattern without escapes is a regular expression: `%` -> `.*`, `_` -> one
byte / one character) and Python's `decimal` arithmetic.  Seeded random strings, patterns and decimals in every
(precision, fraction) shape the mixed-table scenarios use."""
import decimal
import random
import re

import pytest

import emu
import kvfmt
import orc
from tikv_b200 import ffi
from tikv_b200.plan import ColumnDef, Plan, col, const_bytes, const_decimal, const_int, eq, ge, gt, in_, le, like, lt, ne, nulleq

TABLE = 77
COLS = [ColumnDef(100, pk_handle=True), ColumnDef(1, tp=ffi.TP_VARCHAR), ColumnDef(2, tp=ffi.TP_NEWDECIMAL)]
SHAPES = [(1, 0), (9, 0), (10, 2), (14, 4), (18, 9), (30, 10), (65, 30), (5, 5), (20, 0), (12, 11)]


def rand_decimal(rng):
    prec, frac = rng.choice(SHAPES)
    digs = "".join(rng.choice("0123456789") for _ in range(prec)) if rng.random() < 0.85 else ("0" * prec if rng.random() < 0.5 else "0" * (prec - 1) + "1")
    if rng.random() < 0.3:  # small magnitudes collide across shapes: equality cases
        digs = "0" * (prec - 1) + rng.choice("0125")
    txt = (digs[:prec - frac] or "0") + ("." + digs[prec - frac:] if frac else "")
    return ("-" if rng.random() < 0.5 else "") + txt, prec, frac


def build(rng, n, alphabet):
    r = kvfmt.Region()
    rows = []
    for h in range(n):
        s = None if rng.random() < 0.1 else "".join(rng.choice(alphabet) for _ in range(rng.choice([0, 1, 2, 3, 5, 8, 13])))
        d = None if rng.random() < 0.1 else rand_decimal(rng)
        cells = [(1, None if s is None else s.encode(), "bytes"), (2, d, "decimal")]
        r.put(kvfmt.row_key(TABLE, h), kvfmt.row_v2(cells), 10, 20)
        rows.append((h, s, d))
    return r.build(read_ts=100), rows


def like_regex(pattern, binary):
    out = []
    for ch in (pattern.encode() if binary else pattern):
        c = bytes([ch]) if binary else ch
        if c in (b"%", "%"):
            out.append(b".*" if binary else ".*")
        elif c in (b"_", "_"):
            out.append(b"." if binary else ".")
        else:
            out.append(re.escape(c))
    return re.compile((b"" if binary else "").join(out), re.S)


@pytest.mark.parametrize("run", [orc.dag_handle, emu.dag_handle], ids=["oracle", "device-logic"])
def test_like_against_regular_expressions(run):
    rng = random.Random(20260922)
    for collation, alphabet in ((63, "ab%_c"), (-46, "abé测🐶c"), (63, "abé测c")):
        region, rows = build(rng, 300, alphabet)
        binary = collation == 63
        pats = ["%", "_", "", "a%", "%a", "%a%", "_b%", "a_c", "%_%_%", "ab", "%é%", "测_", "__", "%b_", "a%b%c", "___%"]
        pats += ["".join(rng.choice("ab%_c" if binary and "é" not in alphabet else "abé%_测") for _ in range(rng.randrange(1, 6))) for _ in range(14)]
        for i in range(0, len(pats), 10):
            chunk = pats[i:i + 10]
            exprs = [like(col(1, tp=ffi.TP_VARCHAR, collation=collation), const_bytes(p.encode(), collation), escape=0, collation=collation) for p in chunk]  # (escape 0: no character is one)
            plan = Plan().table_scan(TABLE, COLS).projection(col(0), *exprs).build()
            res = run(plan, [kvfmt.table_range(TABLE)], region)
            assert res.status == 0, res.message
            got = {r[0]: r[1:] for r in res.rows()}
            for h, s, _ in rows:
                for p, g in zip(chunk, got[h]):
                    if s is None:
                        assert g is None
                        continue
                    want = like_regex(p, binary).fullmatch(s.encode() if binary else s) is not None
                    assert g == int(want), (collation, s, p, g, want)


@pytest.mark.parametrize("run", [orc.dag_handle, emu.dag_handle], ids=["oracle", "device-logic"])
def test_decimal_comparisons_against_python_decimal(run):
    rng = random.Random(7)
    ctx = decimal.Context(prec=200)
    region, rows = build(rng, 400, "ab")
    dc = col(2, tp=ffi.TP_NEWDECIMAL)
    for _ in range(12):
        cv = rand_decimal(rng)
        k = const_decimal(kvfmt.decimal_bin(*cv))
        other = rand_decimal(rng)
        exprs = [f(dc, k) for f in (lt, le, gt, ge, eq, ne, nulleq)] + [in_(dc, k, const_decimal(kvfmt.decimal_bin(*other)))]
        plan = Plan().table_scan(TABLE, COLS).projection(col(0), *exprs).build()
        res = run(plan, [kvfmt.table_range(TABLE)], region)
        assert res.status == 0, res.message
        got = {r[0]: r[1:] for r in res.rows()}
        c, o = ctx.create_decimal(cv[0]), ctx.create_decimal(other[0])
        for h, _, d in rows:
            g = got[h]
            if d is None:
                assert g[:6] == (None,) * 6 and g[6] == 0 and g[7] is None
                continue
            x = ctx.create_decimal(d[0])
            want = (x < c, x <= c, x > c, x >= c, x == c, x != c, x == c, x == c or x == o)
            assert g == tuple(int(w) for w in want), (d, cv, other, g, want)
    # a selection + count over the same column: the rows the predicate keeps
    cv = ("0.5", 3, 2)
    plan = Plan().table_scan(TABLE, COLS).selection(ge(dc, const_decimal(kvfmt.decimal_bin(*cv)))).aggregation([("count", const_int(1))]).build()
    res = run(plan, [kvfmt.table_range(TABLE)], region)
    assert res.rows() == [(sum(1 for _, _, d in rows if d is not None and ctx.create_decimal(d[0]) >= decimal.Decimal("0.5")),)]


@pytest.mark.parametrize("run", [orc.dag_handle, emu.dag_handle], ids=["oracle", "device-logic"])
def test_time_and_duration_comparisons_against_python_tuples(run):
    """DATETIME(fsp 3) / DATE columns against constants of the other type and fsp, DURATION against nanosecond constants:
    `Ord for Time` is the order of (year, month, day, hour, minute, second, microsecond) — the fsp / type nibble must not
    take part — and `Ord for Duration` the order of the signed nanoseconds."""
    from tikv_b200.plan import const_duration, const_time
    rng = random.Random(99)
    cols = [ColumnDef(100, pk_handle=True), ColumnDef(1, tp=ffi.TP_DATETIME, decimal=3), ColumnDef(2, tp=ffi.TP_DATE), ColumnDef(3, tp=ffi.TP_DURATION)]

    def rand_time(date):
        if rng.random() < 0.08:
            return (0, 0, 0, 0, 0, 0, 0)
        t = (rng.choice([1, 999, 2024, 2024, 2025, 9999]), rng.randrange(1, 13), rng.randrange(1, 29))
        return t + ((0, 0, 0, 0) if date else (rng.randrange(24), rng.randrange(60), rng.randrange(60), rng.randrange(1000) * 1000))
    r = kvfmt.Region()
    rows = []
    for h in range(300):
        dt, d = (None if rng.random() < 0.1 else rand_time(False)), (None if rng.random() < 0.1 else rand_time(True))
        du = None if rng.random() < 0.1 else rng.choice([0, 1, -1, 10 ** 9, -3 * 10 ** 12, rng.randrange(-10 ** 15, 10 ** 15)])
        r.put(kvfmt.row_key(TABLE, h), kvfmt.row_v2([(1, None if dt is None else kvfmt.time_packed(*dt), "time"), (2, None if d is None else kvfmt.time_packed(*d), "time"), (3, du, "duration")]), 10, 20)
        rows.append((h, dt, d, du))
    region = r.build(read_ts=100)
    c_dt, c_d, c_du = col(1, tp=ffi.TP_DATETIME), col(2, tp=ffi.TP_DATE), col(3, tp=ffi.TP_DURATION)
    for _ in range(8):
        kt, kd, kn = rand_time(False), rand_time(True), rng.choice([0, 1, -1, 10 ** 9, rng.randrange(-10 ** 15, 10 ** 15)])
        exprs = [lt(c_dt, const_time(kvfmt.time_packed(*kt))), ge(c_dt, const_time(kvfmt.time_packed(*kd), ffi.TP_DATE)), eq(c_d, const_time(kvfmt.time_packed(*kd), ffi.TP_DATE)),
                 le(c_d, const_time(kvfmt.time_packed(*kt))), gt(c_dt, c_d), nulleq(c_d, c_dt), lt(c_du, const_duration(kn)), ne(c_du, const_duration(kn)),
                 in_(c_d, const_time(kvfmt.time_packed(*kd), ffi.TP_DATE), const_time(0, ffi.TP_DATE))]
        res = run(Plan().table_scan(TABLE, cols).projection(col(0), *exprs).build(), [kvfmt.table_range(TABLE)], region)
        assert res.status == 0, res.message
        got = {x[0]: x[1:] for x in res.rows()}

        def n(v):  # NULL-propagating comparison result
            return None if v is None else int(v)
        for h, dt, d, du in rows:
            want = (n(None if dt is None else dt < kt), n(None if dt is None else dt >= kd), n(None if d is None else d == kd), n(None if d is None else d <= kt),
                    n(None if dt is None or d is None else dt > d), int((d is None and dt is None) or (d is not None and dt is not None and d == dt)),
                    n(None if du is None else du < kn), n(None if du is None else du != kn), n(None if d is None else d in (kd, (0,) * 7)))
            assert got[h] == want, (h, dt, d, du, kt, kd, kn, got[h], want)


@pytest.mark.parametrize("run", [orc.dag_handle, emu.dag_handle], ids=["oracle", "device-logic"])
def test_reference_comparison_truth_table(run):
    """impl_compare.rs generate_numeric_compare_cases (63 rows over {NULL, 3.5, -2.1} x {Gt, Ge, Lt, Le, Eq, Ne, NullEq}), the
    table behind the reference's test_compare_real / test_compare_duration (values mapped through
    Duration::from_millis(v * 1000)) / test_compare_decimal (f64 -> Decimal): tests/golden/compare_cases.json, extracted
    from the reference by tests/golden/gen_compare_cases.py."""
    import json
    import os
    from tikv_b200.plan import const_duration, const_real, null
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "compare_cases.json")))["cases"]
    assert len(cases) == 63
    ops = {"Gt": gt, "Ge": ge, "Lt": lt, "Le": le, "Eq": eq, "Ne": ne, "NullEq": nulleq}
    kinds = {
        "real": lambda v: null(ffi.TP_DOUBLE) if v is None else const_real(v),
        "duration": lambda v: null(ffi.TP_DURATION) if v is None else const_duration(int(v * 1000.0) * 1_000_000),
        "decimal": lambda v: null(ffi.TP_NEWDECIMAL) if v is None else const_decimal(kvfmt.decimal_bin(repr(v), 2, 1)),
    }
    r = kvfmt.Region()
    r.put(kvfmt.row_key(TABLE, 1), kvfmt.row_v2([(1, b"x", "bytes")]), 10, 20)
    region = r.build(read_ts=100)
    for kind, mk in kinds.items():
        for i in range(0, len(cases), 9):
            chunk = cases[i:i + 9]
            exprs = [ops[c["op"]](mk(c["a"]), mk(c["b"])) for c in chunk]
            res = run(Plan().table_scan(TABLE, COLS).projection(*exprs).build(), [kvfmt.table_range(TABLE)], region)
            assert res.status == 0, (kind, res.message)
            assert list(res.rows()[0]) == [c["expect"] for c in chunk], (kind, chunk, res.rows()[0])
