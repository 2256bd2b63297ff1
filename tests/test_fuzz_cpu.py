"""Mutation fuzzing of the device's per-row logic (host build of b2_device.h, tests/host_emul.cpp) against the oracle.

Seeded, deterministic.  Every case takes a small region of well-formed v1 / v2 rows and write records and damages a few
of them at the byte level (truncation, bit flips, duplicated / dropped bytes, trailing garbage, in the row value or in
the write record), then runs scan / selection / aggregation plans through both implementations.  The statuses must agree and,
for a scan, the rows delivered before the first error must agree too (first-error-wins)."""
import random

import pytest

import emu
import kvfmt
import orc
import scenarios as sc
from tikv_b200 import ffi
from tikv_b200.plan import Plan, col, const_int, lt


def _mutate(rng, b):
    b = bytearray(b)
    if not b:
        return bytes(b)
    k = rng.randrange(6)
    i = rng.randrange(len(b))
    if k == 0:
        b = b[:i]                                   # truncate
    elif k == 1:
        b[i] ^= 1 << rng.randrange(8)               # bit flip
    elif k == 2:
        b[i] = rng.choice([0, 1, 3, 4, 5, 8, 9, 0x80, 0xff])  # a datum flag / varint continuation where none belongs
    elif k == 3:
        b = b[:i] + b[i:i + 1] + b[i:]              # duplicated byte
    elif k == 4:
        del b[i]                                    # dropped byte
    else:
        b += bytes(rng.randrange(256) for _ in range(rng.randrange(1, 4)))  # trailing garbage
    return bytes(b)


def _v2_ids_sorted(val):
    """Row v2: are the non-null ids and the null ids each strictly increasing?  (Rows that are not are excluded: the
    reference binary-searches the id arrays (row_slice.rs:175-215), so what it finds in an unsorted array is an artefact
    of Rust's std probe order; the device tries the column's usual position first.  DESIGN.md 3.2.)"""
    if len(val) < 6 or val[0] != 0x80:
        return True
    big = val[1] & 1
    nn, nl = int.from_bytes(val[2:4], "little"), int.from_bytes(val[4:6], "little")
    w = 4 if big else 1
    ids = [int.from_bytes(val[6 + i * w:6 + (i + 1) * w], "little") for i in range(nn + nl) if 6 + (i + 1) * w <= len(val)]
    a, b = ids[:nn], ids[nn:]
    return all(x < y for x, y in zip(a, a[1:])) and all(x < y for x, y in zip(b, b[1:]))


def _short_value(rec):
    """Best-effort: the short value a (possibly damaged) write record would yield (write.rs:296-361)."""
    i = 1
    while i < len(rec) and rec[i] & 0x80:
        i += 1
    i += 1
    if i + 1 < len(rec) and rec[i] == ord("v"):
        return rec[i + 2:i + 2 + rec[i + 1]]
    return b""


def _region(seed, damage):
    rng = random.Random(seed)
    r = kvfmt.Region()
    for h in range(40):
        key = kvfmt.row_key(sc.TABLE, h * 3)
        fmt = rng.choice([1, 2])
        val = sc._row_value(rng, fmt, full_range=False)
        x = rng.random()
        if damage == "value" and x < 0.03:
            m = _mutate(rng, val)
            val = m if _v2_ids_sorted(m) else val
        if damage == "write" and x < 0.02:
            rec = kvfmt.write_record(b"P", 10, short_value=val)
            m = _mutate(rng, rec)
            r.raw_write(key, 20, m if _v2_ids_sorted(_short_value(m)) else rec)
            continue
        if rng.random() < 0.2:
            r.put(key, sc._row_value(rng, fmt, full_range=False), 3, 5)  # an older version below
        r.put(key, val, 10, 20)
        if rng.random() < 0.1:
            r.lock_rec(key, 30, 31)
    return r


PLANS = [("scan", lambda: Plan().table_scan(sc.TABLE, sc.COLUMNS).build()),
         ("subset", lambda: Plan().table_scan(sc.TABLE, sc.COLUMNS).build(output_offsets=[sc.C_H, sc.C6])),
         ("filter", lambda: Plan().table_scan(sc.TABLE, sc.COLUMNS).selection(lt(col(sc.C2), const_int(10))).build(output_offsets=[sc.C_H, sc.C2, sc.C1])),
         ("agg", lambda: Plan().table_scan(sc.TABLE, sc.COLUMNS).aggregation([("count", const_int(1)), ("sum", col(sc.C2))], group_by=[col(sc.C6, tp=ffi.TP_LONG)]).build())]


@pytest.mark.parametrize("damage", ["value", "write"])
def test_mutated_regions_agree_with_oracle(damage):
    n_err = n_ok = 0
    for seed in range(400):
        region = _region(seed * 7 + (1 if damage == "write" else 0), damage).build(read_ts=sc.READ_TS, n_write_blocks=rng_blocks(seed))
        for name, mk in PLANS:
            plan = mk()
            exp = orc.dag_handle(plan, sc.WHOLE, region)
            got = emu.dag_handle(plan, sc.WHOLE, region)
            assert got.status == exp.status, (damage, seed, name, got.status, exp.status, exp.message)
            if exp.status == 0:
                n_ok += 1
                g, e = got.rows(), exp.rows()
                if name == "agg":
                    key = lambda t: tuple((0, 0) if v is None else (1, v) for v in t)
                    g, e = sorted(g, key=key), sorted(e, key=key)
                assert g == e, (damage, seed, name)
            else:
                n_err += 1
                if name != "agg":  # rows before the failing one: the oracle drops its current batch, the device keeps them
                    assert got.rows()[:len(exp.rows())] == exp.rows(), (damage, seed, name)
    assert n_err > 20 and n_ok > 20, (n_err, n_ok)  # the mutations must actually reach both outcomes


def rng_blocks(seed):
    return 1 + seed % 3
