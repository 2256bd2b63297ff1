"""Result comparison helpers: bit-exact for ints / counts / decimals / CRC; f64 sums within a stated tolerance."""
import math
import struct


def _key(row):
    return tuple((0, 0) if v is None else (1, struct.pack("<d", v) if isinstance(v, float) else v) for v in row)


def assert_same_rows(got, exp, ordered=True, float_rel_tol=None, ctx=""):
    assert got.status == exp.status, f"{ctx}: status {got.status} ({getattr(got, 'message', '')}) != oracle {exp.status} ({exp.message})"
    g, e = got.rows(), exp.rows()
    assert len(g) == len(e), f"{ctx}: {len(g)} rows != oracle {len(e)}"
    if not ordered:
        if float_rel_tol is None:
            g, e = sorted(g, key=_key), sorted(e, key=_key)
        else:  # order by the exact (non-float) cells only: float sums may differ in the last bits
            def k2(row):
                return tuple((0, 0) if v is None else (1, v) for v in row if not isinstance(v, float))
            g, e = sorted(g, key=k2), sorted(e, key=k2)
    for i, (a, b) in enumerate(zip(g, e)):
        if float_rel_tol is None:
            ok = _key(a) == _key(b)
        else:
            ok = len(a) == len(b) and all(
                (x is None and y is None) or (x is not None and y is not None and
                 (math.isclose(x, y, rel_tol=float_rel_tol, abs_tol=float_rel_tol) if isinstance(x, float) else x == y))
                for x, y in zip(a, b))
        assert ok, f"{ctx}: row {i}: {a} != oracle {b}"


def assert_topn(got, exp, exact, key_cols, ctx=""):
    """TopN parity: exact rows when ties are broken by the sort key itself; otherwise the multiset of sort keys
    (ties at the cut keep heap-order-dependent rows in the reference)."""
    assert got.status == exp.status, f"{ctx}: status {got.status} ({getattr(got, 'message', '')}) != oracle {exp.status} ({exp.message})"
    g, e = got.rows(), exp.rows()
    assert len(g) == len(e), f"{ctx}: {len(g)} rows != oracle {len(e)}"
    if exact:
        for i, (a, b) in enumerate(zip(g, e)):
            assert _key(a) == _key(b), f"{ctx}: row {i}: {a} != oracle {b}"
    else:
        ka = [tuple(r[c] for c in key_cols) for r in g]
        kb = [tuple(r[c] for c in key_cols) for r in e]
        assert ka == kb, f"{ctx}: sort keys differ"
