"""ctypes wrapper of the CPU oracle (oracle/_build/liborc.so).  TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

from tikv_b200 import ffi
from tikv_b200.plan import key_ranges

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_build", "liborc.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        L = C.CDLL(SO)
        vp = C.c_void_p
        L.orc_dag_handle.argtypes = [C.POINTER(ffi.DagPlan), C.POINTER(ffi.KeyRange), C.c_uint32, C.POINTER(ffi.RegionSource), C.POINTER(vp)]
        L.orc_result_rows.argtypes = [vp]; L.orc_result_rows.restype = C.c_uint64
        L.orc_result_cols.argtypes = [vp]; L.orc_result_cols.restype = C.c_uint32
        L.orc_result_col_kind.argtypes = [vp, C.c_uint32]
        L.orc_result_col_i64.argtypes = [vp, C.c_uint32]; L.orc_result_col_i64.restype = C.POINTER(C.c_int64)
        L.orc_result_col_f64.argtypes = [vp, C.c_uint32]; L.orc_result_col_f64.restype = C.POINTER(C.c_double)
        L.orc_result_col_nonnull.argtypes = [vp, C.c_uint32]; L.orc_result_col_nonnull.restype = C.POINTER(C.c_uint8)
        L.orc_result_col_raw.argtypes = [vp, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.POINTER(C.c_int64))]; L.orc_result_col_raw.restype = C.c_void_p
        L.orc_result_decimal_str.argtypes = [vp, C.c_uint32, C.c_uint64]; L.orc_result_decimal_str.restype = C.c_char_p
        L.orc_result_status.argtypes = [vp]
        L.orc_result_mysql_code.argtypes = [vp]
        L.orc_result_message.argtypes = [vp]; L.orc_result_message.restype = C.c_char_p
        L.orc_result_stats.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.orc_result_free.argtypes = [vp]
        L.orc_result_encoded.argtypes = [vp, C.c_int, C.POINTER(C.c_uint64)]; L.orc_result_encoded.restype = C.POINTER(C.c_uint8)
        L.orc_decimal_write.argtypes = [C.POINTER(ffi.Decimal), C.c_int, C.c_int, C.c_char_p]; L.orc_decimal_write.restype = C.c_size_t
        L.orc_checksum_handle.argtypes = [C.POINTER(ffi.KeyRange), C.c_uint32, C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32,
                                          C.POINTER(ffi.RegionSource), C.POINTER(ffi.ChecksumResponse), C.c_char_p, C.c_size_t]
        L.orc_dag_handle_parallel.argtypes = [C.POINTER(ffi.DagPlan), C.POINTER(ffi.KeyRange), C.c_uint32, C.POINTER(ffi.RegionSource),
                                              C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        L.orc_dag_handle_parallel.restype = C.c_uint64
        L.orc_crc64.argtypes = [C.c_char_p, C.c_size_t]; L.orc_crc64.restype = C.c_uint64
        L.orc_encode_f64_cmp.argtypes = [C.c_double]; L.orc_encode_f64_cmp.restype = C.c_uint64
        L.orc_decode_f64_cmp.argtypes = [C.c_uint64]; L.orc_decode_f64_cmp.restype = C.c_double
        L.orc_encode_i64_cmp.argtypes = [C.c_int64]; L.orc_encode_i64_cmp.restype = C.c_uint64
        for name in ("orc_encode_bytes", "orc_key_append_ts", "orc_encode_var_u64", "orc_encode_var_i64", "orc_decode_var_u64",
                     "orc_decode_var_i64", "orc_encode_row_key", "orc_split_datum", "orc_decimal_to_string"):
            getattr(L, name).restype = C.c_size_t
        L.orc_encode_bytes.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        L.orc_decode_bytes.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.POINTER(C.c_size_t)]; L.orc_decode_bytes.restype = C.c_int64
        L.orc_key_append_ts.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64, C.c_char_p]
        L.orc_encode_var_u64.argtypes = [C.c_uint64, C.c_char_p]
        L.orc_encode_var_i64.argtypes = [C.c_int64, C.c_char_p]
        L.orc_decode_var_u64.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint64)]
        L.orc_decode_var_i64.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int64)]
        L.orc_encode_row_key.argtypes = [C.c_int64, C.c_int64, C.c_char_p]
        L.orc_split_datum.argtypes = [C.c_char_p, C.c_size_t]
        L.orc_write_parse.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint64)]
        L.orc_write_check_gc_fence.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64]
        L.orc_decimal_from_i64.argtypes = [C.c_int64, C.POINTER(ffi.Decimal)]
        L.orc_decimal_from_u64.argtypes = [C.c_uint64, C.POINTER(ffi.Decimal)]
        L.orc_decimal_add.argtypes = [C.POINTER(ffi.Decimal), C.POINTER(ffi.Decimal), C.POINTER(ffi.Decimal)]
        L.orc_decimal_to_string.argtypes = [C.POINTER(ffi.Decimal), C.c_char_p, C.c_size_t]
        L.orc_decimal_cmp.argtypes = [C.POINTER(ffi.Decimal), C.POINTER(ffi.Decimal)]
        _lib = L
    return _lib


class Result:
    """Decoded output of a DAG request: columns as python lists (None = NULL); decimals as python ints."""

    def __init__(self, status, message, mysql_code, columns, kinds, stats):
        self.status, self.message, self.mysql_code = status, message, mysql_code
        self.columns, self.kinds, self.stats = columns, kinds, stats

    @property
    def n_rows(self):
        return len(self.columns[0]) if self.columns else 0

    def rows(self):
        return list(zip(*self.columns)) if self.columns else []


def dag_handle(plan, ranges, region):
    L = lib()
    kr, keep = key_ranges(ranges)
    h = C.c_void_p()
    L.orc_dag_handle(C.byref(plan.c), kr, len(ranges), C.byref(region.c), C.byref(h))
    n = L.orc_result_rows(h)
    cols, kinds = [], []
    for c in range(L.orc_result_cols(h)):
        kind = L.orc_result_col_kind(h, c)
        nn = L.orc_result_col_nonnull(h, c)
        raw_len, offs = C.c_uint64(), C.POINTER(C.c_int64)()
        raw = L.orc_result_col_raw(h, c, C.byref(raw_len), C.byref(offs))
        if raw:  # a column that stayed Raw up to the response: chunk cells (bytes / json / time bits / duration / decimal struct)
            from tikv_b200.executor import raw_cell_values
            o = [offs[i] for i in range(n + 1)] if offs else None
            cols.append(raw_cell_values(kind, C.string_at(raw, raw_len.value), o, [bool(nn[i]) for i in range(n)]))
            kinds.append(kind)
            continue
        if kind == ffi.COL_F64:
            p = L.orc_result_col_f64(h, c)
            vals = [p[i] if nn[i] else None for i in range(n)]
        elif kind == ffi.COL_DECIMAL:
            vals = [int(L.orc_result_decimal_str(h, c, i).decode()) if nn[i] else None for i in range(n)]
        else:
            p = L.orc_result_col_i64(h, c)
            vals = [p[i] if nn[i] else None for i in range(n)]
        cols.append(vals)
        kinds.append(kind)
    st = (C.c_uint64 * 8)()
    L.orc_result_stats(h, st)
    stats = dict(write_next=st[0], write_seek=st[1], over_seek_bound=st[2], processed_keys=st[3], processed_size=st[4],
                 data_processed_keys=st[5], lock_processed_keys=st[6], met_newer=C.c_int64(st[7]).value)
    res = Result(L.orc_result_status(h), L.orc_result_message(h).decode(), L.orc_result_mysql_code(h), cols, kinds, stats)
    L.orc_result_warning_count.argtypes = [C.c_void_p]
    L.orc_result_warning_count.restype = C.c_uint64
    res.warning_count = L.orc_result_warning_count(h)  # SelectResponse.warning_count ("Division by 0", expr/ctx.rs:267-286)
    res.encoded = {}
    for t in (0, 1):  # EncodeType::TypeDefault / TypeChunk
        ln = C.c_uint64()
        pp = L.orc_result_encoded(h, t, C.byref(ln))
        res.encoded[t] = C.string_at(pp, ln.value) if ln.value else b""
    L.orc_result_free(h)
    return res


def checksum(ranges, region, old_prefix=b"", new_prefix=b""):
    L = lib()
    kr, keep = key_ranges(ranges)
    out = ffi.ChecksumResponse()
    err = C.create_string_buffer(256)
    st = L.orc_checksum_handle(kr, len(ranges), old_prefix, len(old_prefix), new_prefix, len(new_prefix),
                               C.byref(region.c), C.byref(out), err, 256)
    return st, (out.checksum, out.total_kvs, out.total_bytes), err.value.decode()


def mvcc_scan(region, lower=None, upper=None, desc=False):
    """Raw forward (or, desc=True, backward) scan (encoded user-key bounds).  Returns (status, [(user_key, value)], stats)."""
    L = lib()
    L.orc_mvcc_scan_backward.argtypes = [C.POINTER(ffi.RegionSource), C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    L.orc_mvcc_scan_backward.restype = C.c_void_p
    L.orc_scan_stats_backward.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.orc_mvcc_scan.argtypes = [C.POINTER(ffi.RegionSource), C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    L.orc_mvcc_scan.restype = C.c_void_p
    L.orc_scan_rows.argtypes = [C.c_void_p]; L.orc_scan_rows.restype = C.c_uint64
    for f in (L.orc_scan_key, L.orc_scan_val):
        f.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_size_t)]
        f.restype = C.POINTER(C.c_uint8)
    L.orc_scan_status.argtypes = [C.c_void_p]
    L.orc_scan_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.orc_scan_free.argtypes = [C.c_void_p]
    h = (L.orc_mvcc_scan_backward if desc else L.orc_mvcc_scan)(C.byref(region.c), lower, len(lower) if lower else 0, upper, len(upper) if upper else 0)
    out = []
    ln = C.c_size_t()
    for i in range(L.orc_scan_rows(h)):
        kp = L.orc_scan_key(h, i, C.byref(ln)); k = bytes(kp[:ln.value])
        vp = L.orc_scan_val(h, i, C.byref(ln)); v = bytes(vp[:ln.value])
        out.append((k, v))
    st = (C.c_uint64 * 8)()
    L.orc_scan_stats(h, st)
    stats = dict(write_next=st[0], write_seek=st[1], over_seek_bound=st[2], processed_keys=st[3], processed_size=st[4],
                 data_processed_keys=st[5], lock_processed_keys=st[6], met_newer=C.c_int64(st[7]).value)
    st2 = (C.c_uint64 * 2)()
    L.orc_scan_stats_backward(h, st2)
    stats.update(write_prev=st2[0], write_seek_for_prev=st2[1])
    status = L.orc_scan_status(h)
    L.orc_scan_free(h)
    return status, out, stats
