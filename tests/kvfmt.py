"""Independent pure-Python encoders for TiKV's on-disk formats, used to build test regions.

Written from the format description, NOT from the oracle's C++ code, so that
oracle and encoders check each other; both are pinned by the reference's golden vectors in
tests/test_oracle_golden.py.
"""
import ctypes as C
import struct

import numpy as np

from tikv_b200 import ffi

SIGN = 1 << 63
U64 = (1 << 64) - 1


def enc_i64_cmp(v):
    return struct.pack(">Q", (v & U64) ^ SIGN)


def enc_u64_desc(v):
    return struct.pack(">Q", (~v) & U64)


def enc_f64_cmp(f):
    (u,) = struct.unpack(">Q", struct.pack(">d", f))
    u = (u | SIGN) if u < SIGN else (~u & U64)
    return struct.pack(">Q", u)


def enc_var_u64(v):
    out = bytearray()
    v &= U64
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def enc_var_i64(v):
    uv = (v << 1) & U64
    if v < 0:
        uv = ~uv & U64
    return enc_var_u64(uv)


def enc_bytes_memcmp(b):
    out = bytearray()
    i = 0
    while True:
        chunk = b[i:i + 8]
        pad = 8 - len(chunk)
        out += chunk + b"\x00" * pad
        out.append(0xFF - pad)
        i += 8
        if pad > 0:
            return bytes(out)


def row_key(table_id, handle):
    return b"t" + enc_i64_cmp(table_id) + b"_r" + enc_i64_cmp(handle)


def write_key(raw_key, commit_ts):
    return enc_bytes_memcmp(raw_key) + enc_u64_desc(commit_ts)


def default_key(raw_key, start_ts):
    return enc_bytes_memcmp(raw_key) + enc_u64_desc(start_ts)


def write_record(wtype, start_ts, short_value=None, overlapped_rollback=False, gc_fence=None, last_change=None,
                 txn_source=0, tail=b""):
    """wtype in b'PDLR'.  last_change = (ts, versions) (ts 0, versions 1 = NotExist)."""
    out = bytearray(wtype) + enc_var_u64(start_ts)
    if short_value is not None:
        assert len(short_value) <= 255
        out += b"v" + bytes([len(short_value)]) + short_value
    if overlapped_rollback:
        out += b"R"
    if gc_fence is not None:
        out += b"F" + struct.pack(">Q", gc_fence)
    if last_change is not None:
        out += b"l" + struct.pack(">Q", last_change[0]) + enc_var_u64(last_change[1])
    if txn_source:
        out += b"S" + enc_var_u64(txn_source)
    return bytes(out) + tail


def lock_record(ltype, primary, ts, ttl=0, short_value=None, for_update_ts=0, txn_size=0, min_commit_ts=0,
                async_secondaries=None):
    out = bytearray(ltype) + enc_var_i64(len(primary)) + primary + enc_var_u64(ts) + enc_var_u64(ttl)
    if short_value is not None:
        out += b"v" + bytes([len(short_value)]) + short_value
    if for_update_ts:
        out += b"f" + struct.pack(">Q", for_update_ts)
    if txn_size:
        out += b"t" + struct.pack(">Q", txn_size)
    if min_commit_ts:
        out += b"c" + struct.pack(">Q", min_commit_ts)
    if async_secondaries is not None:
        out += b"a" + enc_var_u64(len(async_secondaries))
        for s in async_secondaries:
            out += enc_var_i64(len(s)) + s
    return bytes(out)


# ---- datums (v1 rows) ----
NIL, BYTES, COMPACT_BYTES, INT, UINT, FLOAT, DECIMAL, DURATION, VAR_INT, VAR_UINT, JSON = range(11)


def datum_int(v, comparable=False):
    return bytes([INT]) + enc_i64_cmp(v) if comparable else bytes([VAR_INT]) + enc_var_i64(v)


def datum_uint(v, comparable=False):
    return bytes([UINT]) + struct.pack(">Q", v) if comparable else bytes([VAR_UINT]) + enc_var_u64(v)


def datum_f64(f):
    return bytes([FLOAT]) + enc_f64_cmp(f)


def datum_null():
    return bytes([NIL])


def datum_bytes(b):
    return bytes([COMPACT_BYTES]) + enc_var_i64(len(b)) + b


# ---- non Int/Real cells (codec/mysql/{decimal,time,duration,json}) ----
_DIG2BYTES = [0, 1, 1, 2, 2, 3, 3, 4, 4, 4]


def decimal_bin(value, prec, frac):
    """decimal.rs write_decimal (:2025-2132) for a value that fits (prec, frac): [prec, frac, MySQL binary decimal].
    `value` is a decimal.Decimal / int / str."""
    import decimal
    ctx = decimal.Context(prec=200)
    d = ctx.create_decimal(value)
    neg = d < 0
    q = d.copy_abs().quantize(decimal.Decimal(1).scaleb(-frac, ctx), context=ctx) if frac else d.copy_abs().to_integral_value(context=ctx)
    digits = f"{q:f}"
    ip, _, fp = digits.partition(".")
    ip = ip.lstrip("0")
    int_cnt = prec - frac
    assert len(ip) <= int_cnt and len(fp) == frac, (value, prec, frac)
    ip = ip.rjust(int_cnt, "0")
    out = bytearray()
    lead = int_cnt % 9
    if lead:
        out += int(ip[:lead]).to_bytes(_DIG2BYTES[lead], "big")
    for i in range(lead, int_cnt, 9):
        out += int(ip[i:i + 9]).to_bytes(4, "big")
    full = frac // 9 * 9
    for i in range(0, full, 9):
        out += int(fp[i:i + 9]).to_bytes(4, "big")
    if frac % 9:
        out += int(fp[full:]).to_bytes(_DIG2BYTES[frac % 9], "big")
    if neg and any(out):
        out = bytearray(b ^ 0xFF for b in out)
    out[0] ^= 0x80
    return bytes([prec, frac]) + bytes(out)


def time_packed(year, month, day, hour=0, minute=0, second=0, micro=0):
    """Time::to_packed_u64 (mysql/time/mod.rs:2045-2074)."""
    ymd = ((year * 13 + month) << 5) | day
    hms = (hour << 12) | (minute << 6) | second
    return (((ymd << 17) | hms) << 24) | micro


def time_bits(year, month, day, hour=0, minute=0, second=0, micro=0, fsp=0, date=False):
    """The CoreTime bit field of a chunk cell (mysql/time/mod.rs:167-196), written independently of the decoders."""
    return (year << 50) | (month << 46) | (day << 41) | (hour << 36) | (minute << 30) | (second << 24) | (micro << 4) | (0b1110 if date else fsp << 1)


def json_string(sv):
    """binary JSON string: type code 0x0c, varint length, bytes (mysql/json/binary.rs)."""
    b = sv.encode()
    return bytes([0x0C]) + enc_var_u64(len(b)) + b


def json_i64(v):
    return bytes([0x09]) + struct.pack("<q", v)


def datum_decimal(value, prec, frac):
    return bytes([6]) + decimal_bin(value, prec, frac)


def datum_time(packed, comparable=False):
    return datum_uint(packed, comparable)


def datum_duration(nanos, fixed=False):
    return bytes([7]) + enc_i64_cmp(nanos) if fixed else bytes([VAR_INT]) + enc_var_i64(nanos)


def row_v1(cols):
    """cols: list of (col_id, datum_bytes)."""
    out = bytearray()
    for cid, d in cols:
        out += bytes([VAR_INT]) + enc_var_i64(cid) + d
    return bytes(out) if out else bytes([NIL])


def _v2_int(v, unsigned):
    if unsigned:
        v &= U64
        for w, fmt in ((1, "<B"), (2, "<H"), (4, "<I")):
            if v < (1 << (8 * w)):
                return struct.pack(fmt, v)
        return struct.pack("<Q", v)
    for w, fmt in ((1, "<b"), (2, "<h"), (4, "<i")):
        if -(1 << (8 * w - 1)) <= v < (1 << (8 * w - 1)):
            return struct.pack(fmt, v)
    return struct.pack("<q", v)


def row_v2(cols, checksum=None):
    """cols: list of (col_id, value, kind) with kind in {'int','uint','f64','bytes','null'}; value None = NULL.
    Layout per row/v2/row_slice.rs:14-31 and encoder_for_test.rs:353-421."""
    non_null = sorted([(cid, v, k) for cid, v, k in cols if v is not None and k != "null"], key=lambda x: x[0])
    nulls = sorted([cid for cid, v, k in cols if v is None or k == "null"])
    vals = []
    for cid, v, k in non_null:
        if k == "int":
            vals.append(_v2_int(v, False))
        elif k == "uint":
            vals.append(_v2_int(v, True))
        elif k == "f64":
            vals.append(enc_f64_cmp(v))
        elif k == "bytes":
            vals.append(bytes(v))
        elif k == "time":      # packed u64, compact unsigned (compat_v1.rs:83-90)
            vals.append(_v2_int(v, True))
        elif k == "duration":  # nanoseconds, compact signed (:94-100)
            vals.append(_v2_int(v, False))
        elif k == "decimal":   # (value, prec, frac) -> [prec, frac, bin] (:101-105)
            vals.append(decimal_bin(*v))
        else:
            raise ValueError(k)
    total = sum(len(x) for x in vals)
    big = any(cid > 255 for cid, _, _ in non_null) or any(c > 255 for c in nulls) or total > 0xFFFF
    flags = (1 if big else 0) | (2 if checksum is not None else 0)
    out = bytearray([128, flags]) + struct.pack("<HH", len(non_null), len(nulls))
    idf, offf = ("<I", "<I") if big else ("<B", "<H")
    for cid, _, _ in non_null:
        out += struct.pack(idf, cid)
    for cid in nulls:
        out += struct.pack(idf, cid)
    end = 0
    for x in vals:
        end += len(x)
        out += struct.pack(offf, end)
    for x in vals:
        out += x
    if checksum is not None:
        out += bytes([0]) + struct.pack("<I", checksum)
    return bytes(out)


# ---- region assembly ----
class HostBlock:
    """One CF block in host memory (numpy-backed) + its ctypes descriptor."""

    def __init__(self, kvs):
        kvs = sorted(kvs, key=lambda kv: kv[0])
        self.n = len(kvs)
        ko, vo = [0], [0]
        for k, v in kvs:
            ko.append(ko[-1] + len(k))
            vo.append(vo[-1] + len(v))

        def heap(parts, total):
            buf = np.zeros(((total + 15) // 16 + 1) * 16, dtype=np.uint8)
            if total:
                buf[:total] = np.frombuffer(b"".join(parts), dtype=np.uint8)
            return buf

        self.keys = heap([k for k, _ in kvs], ko[-1])
        self.vals = heap([v for _, v in kvs], vo[-1])
        self.key_offs = np.asarray(ko, dtype=np.uint32)
        self.val_offs = np.asarray(vo, dtype=np.uint32)
        self.kvs = kvs
        self.c = ffi.CfBlock()
        self.c.keys, self.c.key_offs = self.keys.ctypes.data, self.key_offs.ctypes.data
        self.c.vals, self.c.val_offs = self.vals.ctypes.data, self.val_offs.ctypes.data
        self.c.n = self.n


class Region:
    """Builds CF_WRITE / CF_DEFAULT / CF_LOCK for one region, the way must_prewrite_put / must_commit would."""

    def __init__(self):
        self.write, self.dflt, self.lock = [], [], []

    def put(self, raw_key, value, start_ts, commit_ts, force_long=False, **kw):
        if len(value) <= 255 and not force_long:
            self.write.append((write_key(raw_key, commit_ts), write_record(b"P", start_ts, short_value=value, **kw)))
        else:
            self.write.append((write_key(raw_key, commit_ts), write_record(b"P", start_ts, **kw)))
            self.dflt.append((default_key(raw_key, start_ts), value))
        return self

    def delete(self, raw_key, start_ts, commit_ts, **kw):
        self.write.append((write_key(raw_key, commit_ts), write_record(b"D", start_ts, **kw)))
        return self

    def lock_rec(self, raw_key, start_ts, commit_ts, **kw):
        self.write.append((write_key(raw_key, commit_ts), write_record(b"L", start_ts, **kw)))
        return self

    def rollback(self, raw_key, start_ts, **kw):
        self.write.append((write_key(raw_key, start_ts), write_record(b"R", start_ts, **kw)))
        return self

    def raw_write(self, raw_key, commit_ts, value_bytes):
        self.write.append((write_key(raw_key, commit_ts), value_bytes))
        return self

    def add_lock(self, raw_key, lock_bytes):
        self.lock.append((enc_bytes_memcmp(raw_key), lock_bytes))
        return self

    def build(self, read_ts, n_write_blocks=1, isolation=ffi.ISO_SI, check_newer=True, bypass=(), access=()):
        return RegionHost(self, read_ts, n_write_blocks, isolation, check_newer, bypass, access)


class RegionHost:
    def __init__(self, r, read_ts, n_write_blocks, isolation, check_newer, bypass, access):
        w = sorted(r.write, key=lambda kv: kv[0])
        # split into blocks at user-key boundaries (all versions of a key stay in one block)
        self.wblocks = []
        if n_write_blocks <= 1 or len(w) < 2:
            self.wblocks.append(HostBlock(w))
        else:
            per = max(1, len(w) // n_write_blocks)
            start = 0
            while start < len(w):
                end = min(len(w), start + per)
                while end < len(w) and w[end][0][:-8] == w[end - 1][0][:-8]:
                    end += 1
                self.wblocks.append(HostBlock(w[start:end]))
                start = end
        self.dblock = HostBlock(r.dflt) if r.dflt else None
        self.lblock = HostBlock(r.lock) if r.lock else None
        self.n_entries = len(w)
        self._warr = (ffi.CfBlock * len(self.wblocks))(*[b.c for b in self.wblocks])
        self._bypass = (C.c_uint64 * max(1, len(bypass)))(*bypass)
        self._access = (C.c_uint64 * max(1, len(access)))(*access)
        s = ffi.RegionSource()
        s.location, s.device = ffi.LOC_HOST, 0
        s.write, s.n_write = self._warr, len(self.wblocks)
        if self.dblock:
            self._darr = (ffi.CfBlock * 1)(self.dblock.c)
            s.dflt, s.n_dflt = self._darr, 1
        if self.lblock:
            self._larr = (ffi.CfBlock * 1)(self.lblock.c)
            s.lock = self._larr
        s.read_ts, s.isolation_level, s.check_has_newer_ts_data = read_ts, isolation, int(check_newer)
        s.bypass_locks, s.n_bypass_locks = self._bypass, len(bypass)
        s.access_locks, s.n_access_locks = self._access, len(access)
        self.c = s


def table_range(table_id, lo=None, hi=None):
    """Raw key range covering handles [lo, hi) of a table (whole table if None)."""
    start = row_key(table_id, lo) if lo is not None else b"t" + enc_i64_cmp(table_id) + b"_r"
    end = row_key(table_id, hi) if hi is not None else b"t" + enc_i64_cmp(table_id) + b"_s"
    return (start, end)


# ---- response decoders (independent of both the oracle and the device encoders) ------------------------------------
def _dec_bin_to_int(buf, pos, prec, frac):
    """MySQL binary decimal -> (python int scaled by 10**frac, new pos)."""
    d2b = [0, 1, 1, 2, 2, 3, 3, 4, 4, 4]
    ints, fr = prec - frac, frac
    size = (ints // 9) * 4 + d2b[ints % 9] + (fr // 9) * 4 + d2b[fr % 9]
    raw = bytearray(buf[pos:pos + size])
    neg = not (raw[0] & 0x80)
    raw[0] ^= 0x80
    if neg:
        raw = bytearray(b ^ 0xFF for b in raw)
    p, digits = 0, ""
    for cnt in ([ints % 9] if ints % 9 else []) + [9] * (ints // 9) + [9] * (fr // 9) + ([fr % 9] if fr % 9 else []):
        nb = d2b[cnt]
        digits += str(int.from_bytes(raw[p:p + nb], "big")).rjust(cnt, "0")
        p += nb
    v = int(digits or "0")
    return (-v if neg else v), pos + size


def decode_datum(buf, pos):
    """One datum of a TypeDefault row -> (python value, new pos).  Decimals come back as (unscaled int, frac)."""
    flag = buf[pos]
    pos += 1
    if flag == 0:
        return None, pos
    if flag in (3, 4):
        u = int.from_bytes(buf[pos:pos + 8], "big")
        if flag == 3:
            u ^= 1 << 63
            if u >= 1 << 63:
                u -= 1 << 64
        return u, pos + 8
    if flag == 5:
        u = int.from_bytes(buf[pos:pos + 8], "big")
        u = u ^ (1 << 63) if u & (1 << 63) else u ^ ((1 << 64) - 1)
        return struct.unpack("<d", struct.pack("<Q", u))[0], pos + 8
    if flag == 6:
        prec, frac = buf[pos], buf[pos + 1]
        v, pos = _dec_bin_to_int(buf, pos + 2, prec, frac)
        return (v, frac), pos
    if flag in (8, 9):
        u, shift = 0, 0
        while True:
            b = buf[pos]
            pos += 1
            u |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                break
        return ((u >> 1) ^ -(u & 1) if flag == 8 else u), pos
    raise ValueError(f"datum flag {flag}")


def decode_datum_rows(buf, n_cols):
    rows, pos = [], 0
    while pos < len(buf):
        row = []
        for _ in range(n_cols):
            v, pos = decode_datum(buf, pos)
            row.append(v)
        rows.append(tuple(row))
    return rows


def decode_chunk(buf, elem_sizes):
    """TypeChunk column blocks -> list of columns of raw little-endian cell bytes (None = NULL)."""
    cols, pos = [], 0
    for es in elem_sizes:
        n, nulls = struct.unpack_from("<II", buf, pos)
        pos += 8
        bm = None
        if nulls:
            bm = buf[pos:pos + (n + 7) // 8]
            pos += (n + 7) // 8
        cells = []
        for i in range(n):
            cell = bytes(buf[pos + i * es:pos + (i + 1) * es])
            cells.append(cell if (bm is None or (bm[i >> 3] >> (i & 7)) & 1) else None)
        pos += n * es
        cols.append(cells)
    assert pos == len(buf), (pos, len(buf))
    return cols


def decimal_struct_value(cell):
    """40-byte MyDecimal struct (decimal.rs:927-942) -> (unscaled python int, frac digits).  The struct is not canonical
    (digitsInt may include leading zero words left behind by the order of additions), the value is."""
    int_cnt, frac_cnt, _res, neg = cell[0], cell[1], cell[2], cell[3]
    words = struct.unpack_from("<9I", cell, 4)
    iw, fw = (int_cnt + 8) // 9, (frac_cnt + 8) // 9
    v = 0
    for w in words[:iw]:
        v = v * 10 ** 9 + w
    for w in words[iw:iw + fw]:
        v = v * 10 ** 9 + w
    v //= 10 ** (fw * 9 - frac_cnt)
    return (-v if neg else v), frac_cnt


def dec_bytes_memcmp(enc):
    """Inverse of enc_bytes_memcmp (tikv_util/src/codec/bytes.rs:178-228)."""
    out, pos = b"", 0
    while True:
        grp, marker = enc[pos:pos + 8], enc[pos + 8]
        pad = 0xFF - marker
        out += grp[:8 - pad]
        pos += 9
        if pad:
            return out


def index_key(table_id, index_id, payload=b""):
    """t{table_id}_i{index_id}{payload}: payload = the index columns as memcomparable datums (+ the handle datum for a
    non-unique index) — table.rs encode_index_seek_key."""
    return b"t" + enc_i64_cmp(table_id) + b"_i" + enc_i64_cmp(index_id) + payload
