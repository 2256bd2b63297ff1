"""ctypes wrapper of tests/host_emul.cpp: the device row logic driven on the CPU (debug harness, not product)."""
import ctypes as C
import os
import subprocess

from tikv_b200 import ffi
from tikv_b200.executor import _decimal_to_int
from tikv_b200.plan import key_ranges

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_emul.cpp")
SO = os.path.join(ROOT, "tests", "_build", "libemu.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        deps = [SRC] + [os.path.join(ROOT, "tikv_b200", "csrc", f) for f in ("b2_device.h", "plan_compile.h")]
        if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
            os.makedirs(os.path.dirname(SO), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SO, SRC])
        L = C.CDLL(SO)
        vp = C.c_void_p
        L.emu_dag_handle.argtypes = [C.POINTER(ffi.DagPlan), C.POINTER(ffi.KeyRange), C.c_uint32, C.POINTER(ffi.RegionSource)]
        L.emu_dag_handle.restype = vp
        L.emu_checksum.argtypes = [C.POINTER(ffi.KeyRange), C.c_uint32, C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.POINTER(ffi.RegionSource)]
        L.emu_checksum.restype = vp
        for f, rt in (("emu_status", C.c_int), ("emu_dev_err", C.c_int), ("emu_err_entry", C.c_uint64), ("emu_message", C.c_char_p),
                      ("emu_rows", C.c_uint64), ("emu_cols", C.c_uint32)):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = rt
        L.emu_col_kind.argtypes = [vp, C.c_uint32]
        L.emu_col_data.argtypes = [vp, C.c_uint32]; L.emu_col_data.restype = C.POINTER(C.c_uint64)
        L.emu_col_dec.argtypes = [vp, C.c_uint32]; L.emu_col_dec.restype = C.POINTER(C.c_uint8)
        L.emu_col_nonnull.argtypes = [vp, C.c_uint32]; L.emu_col_nonnull.restype = C.POINTER(C.c_uint8)
        L.emu_stats.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.emu_free.argtypes = [vp]
        L.emu_check_supported.argtypes = [C.POINTER(ffi.DagPlan), C.c_char_p, C.c_size_t]
        L.emu_set_fast_front.argtypes = [C.c_int]
        L.emu_parse_decimal.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p]
        L.emu_fast_hits.restype = C.c_uint64
        _lib = L
    return _lib


class EmuResult:
    def __init__(self, status, message, dev_err, err_entry, columns, kinds, stats):
        self.status, self.message, self.dev_err, self.err_entry = status, message, dev_err, err_entry
        self.columns, self.kinds, self.stats = columns, kinds, stats

    @property
    def n_rows(self):
        return len(self.columns[0]) if self.columns else 0

    def rows(self):
        return list(zip(*self.columns)) if self.columns else []


def _stats(L, h):
    st = (C.c_uint64 * 7)()
    L.emu_stats(h, st)
    return dict(processed_keys=st[0], processed_size=st[1], met_newer=st[2], default_lookups=st[3], checksum=st[4], total_kvs=st[5], total_bytes=st[6])


def plan_scan_only(plan):
    """No aggregation in the plan: a DECIMAL output column is a stored column (cell references), not a SUM result."""
    return all(plan.c.executors[i].tp not in (ffi.EXEC_AGGREGATION, ffi.EXEC_STREAM_AGG) for i in range(plan.c.n_executors))


def dag_handle(plan, ranges, region):
    import struct
    L = lib()
    kr, keep = key_ranges(ranges)
    h = L.emu_dag_handle(C.byref(plan.c), kr, len(ranges), C.byref(region.c))
    n = L.emu_rows(h)
    cols, kinds = [], []
    for c in range(L.emu_cols(h)):
        kind = L.emu_col_kind(h, c)
        nn = L.emu_col_nonnull(h, c)
        d = L.emu_col_data(h, c)
        if kind in (ffi.COL_BYTES, ffi.COL_JSON) or (kind == ffi.COL_DECIMAL and plan_scan_only(plan)):
            # the kernels leave cell references (address << 16 | length); the raw_* kernels of the device path resolve them
            from tikv_b200.executor import _decimal_value
            vals = []
            for i in range(n):
                if not nn[i]:
                    vals.append(None)
                    continue
                cell = C.string_at(d[i] >> 16, d[i] & 0xffff)
                if kind == ffi.COL_DECIMAL:
                    out = (C.c_uint8 * 40)()
                    assert L.emu_parse_decimal(cell, len(cell), out) == 1, cell
                    vals.append(_decimal_value(bytes(out)))
                else:
                    vals.append(cell)
        elif kind == ffi.COL_TIME:
            vals = [d[i] if nn[i] else None for i in range(n)]
        elif kind == ffi.COL_F64:
            vals = [struct.unpack("<d", struct.pack("<Q", d[i]))[0] if nn[i] else None for i in range(n)]
        elif kind == ffi.COL_DECIMAL:
            raw = L.emu_col_dec(h, c)
            vals = [_decimal_to_int(bytes(raw[40 * i:40 * i + 40])) if nn[i] else None for i in range(n)]
        else:
            vals = [C.c_int64(d[i]).value if nn[i] else None for i in range(n)]
        cols.append(vals)
        kinds.append(kind)
    res = EmuResult(L.emu_status(h), L.emu_message(h).decode(), L.emu_dev_err(h), L.emu_err_entry(h), cols, kinds, _stats(L, h))
    L.emu_free(h)
    return res


def checksum(ranges, region, old_prefix=b"", new_prefix=b""):
    L = lib()
    kr, keep = key_ranges(ranges)
    h = L.emu_checksum(kr, len(ranges), old_prefix, len(old_prefix), new_prefix, len(new_prefix), C.byref(region.c))
    st = _stats(L, h)
    status = L.emu_status(h)
    L.emu_free(h)
    return status, (st["checksum"], st["total_kvs"], st["total_bytes"])


def check_supported(plan):
    msg = C.create_string_buffer(256)
    rc = lib().emu_check_supported(C.byref(plan.c), msg, 256)
    return rc, msg.value.decode()


def set_fast_front(on):
    """Switch the clean-entry front end (entry_fast) of the emulation on / off; returns nothing."""
    lib().emu_set_fast_front(1 if on else 0)


def fast_hits():
    """Entries the clean-entry front end has accepted so far (process-wide counter)."""
    return lib().emu_fast_hits()
