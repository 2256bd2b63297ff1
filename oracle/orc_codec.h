// TEST INFRASTRUCTURE — CPU oracle (restatement of the reference algorithm), not product code.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference may use it.
//
// Number / bytes / key codecs.  Each function cites the tikv/tikv file:line it follows.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace orc {

typedef std::vector<uint8_t> Bytes;

struct Slice {
  const uint8_t* p = nullptr;
  size_t n = 0;
  Slice() {}
  Slice(const uint8_t* p_, size_t n_) : p(p_), n(n_) {}
  bool empty() const { return n == 0; }
  uint8_t operator[](size_t i) const { return p[i]; }
  Slice sub(size_t from) const { return Slice(p + from, n - from); }
  Slice sub(size_t from, size_t len) const { return Slice(p + from, len); }
};

inline int cmp_bytes(Slice a, Slice b) {
  size_t m = a.n < b.n ? a.n : b.n;
  int c = m ? memcmp(a.p, b.p, m) : 0;
  if (c) return c;
  return a.n < b.n ? -1 : (a.n > b.n ? 1 : 0);
}

// ---- fixed-width numbers: components/tikv_util/src/codec/number.rs:12-75 ----
const uint64_t SIGN_MARK = 0x8000000000000000ull;

inline void put_u64_be(Bytes& b, uint64_t v) {
  for (int i = 7; i >= 0; --i) b.push_back((uint8_t)(v >> (8 * i)));
}
inline uint64_t get_u64_be(const uint8_t* p) {
  uint64_t v = 0;
  for (int i = 0; i < 8; ++i) v = (v << 8) | p[i];
  return v;
}
inline uint64_t get_u64_le(const uint8_t* p) {
  uint64_t v = 0;
  for (int i = 7; i >= 0; --i) v = (v << 8) | p[i];
  return v;
}
// number.rs:19-25 encode_i64 = (v as u64) ^ SIGN_MARK, big-endian
inline void encode_i64(Bytes& b, int64_t v) { put_u64_be(b, (uint64_t)v ^ SIGN_MARK); }
inline int64_t decode_i64(const uint8_t* p) { return (int64_t)(get_u64_be(p) ^ SIGN_MARK); }
// number.rs:72-75 encode_u64_desc = !v big-endian
inline void encode_u64_desc(Bytes& b, uint64_t v) { put_u64_be(b, ~v); }
inline uint64_t decode_u64_desc(const uint8_t* p) { return ~get_u64_be(p); }
// number.rs:27-42 order-preserving f64
inline uint64_t encode_f64_to_cmp_u64(double f) {
  uint64_t u;
  memcpy(&u, &f, 8);
  if ((int64_t)u >= 0) u |= SIGN_MARK; else u = ~u;
  return u;
}
inline double decode_cmp_u64_to_f64(uint64_t u) {
  if (u & SIGN_MARK) u &= ~SIGN_MARK; else u = ~u;
  double f;
  memcpy(&f, &u, 8);
  return f;
}

// ---- varint: components/codec/src/number.rs:417-525 (LEB128; zig-zag for signed) ----
inline void encode_var_u64(Bytes& b, uint64_t v) {
  while (v >= 0x80) { b.push_back((uint8_t)(v | 0x80)); v >>= 7; }
  b.push_back((uint8_t)v);
}
// components/codec/src/number.rs:445-483 try_decode_var_u64 (the decoder behind read_var_u64 /
// read_var_i64 / datum payloads).  Returns bytes consumed, 0 on eof.  With >= 10 bytes available the
// 10th byte contributes one bit and always terminates (no overflow error), exactly as the reference.
inline size_t decode_var_u64(Slice s, uint64_t* out) {
  uint64_t v = 0;
  if (s.n >= 10) {
    for (size_t i = 0; i < 9; ++i) {
      uint64_t b = s[i];
      v |= (b & 0x7f) << (7 * i);
      if (b < 0x80) { *out = v; return i + 1; }
    }
    v |= ((uint64_t)s[9] & 1) << 63;
    *out = v;
    return 10;
  }
  size_t i = 0;
  unsigned shift = 0;
  while (i < s.n && s[i] >= 0x80) { v |= (uint64_t)(s[i] & 0x7f) << shift; shift += 7; ++i; }
  if (i == s.n) return 0;
  v |= (uint64_t)s[i] << shift;
  *out = v;
  return i + 1;
}
// components/tikv_util/src/codec/number.rs:224-275 decode_var_u64 (used by WriteRef::parse for the
// last_change / txn_source fields and by Lock::parse): 10th byte > 1 is an overflow error.
inline size_t decode_var_u64_tu(Slice s, uint64_t* out) {
  uint64_t v = 0;
  for (size_t i = 0; i < s.n && i < 10; ++i) {
    uint64_t b = s[i];
    if (i == 9) {
      if (b > 1) return 0;
      v |= b << 63;
      *out = v;
      return 10;
    }
    v |= (b & 0x7f) << (7 * i);
    if (b < 0x80) { *out = v; return i + 1; }
  }
  return 0;
}
inline void encode_var_i64(Bytes& b, int64_t v) {  // number.rs:496-502
  uint64_t uv = (uint64_t)v << 1;
  if (v < 0) uv = ~uv;
  encode_var_u64(b, uv);
}
inline size_t decode_var_i64(Slice s, int64_t* out) {  // number.rs:515-525
  uint64_t uv;
  size_t n = decode_var_u64(s, &uv);
  if (!n) return 0;
  int64_t v = (int64_t)(uv >> 1);
  if (uv & 1) v = ~v;
  *out = v;
  return n;
}
// number.rs:530-567 get_first_encoded_var_int_len
inline size_t first_var_int_len(Slice s) {
  if (s.n >= 10) {
    for (size_t i = 0; i < 9; ++i)
      if (s[i] < 0x80) return i + 1;
    return 10;
  }
  for (size_t i = 0; i < s.n; ++i)
    if (s[i] < 0x80) return i + 1;
  return s.n;
}

// ---- memcomparable bytes: components/tikv_util/src/codec/bytes.rs:13-55, 178-228 ----
const size_t ENC_GROUP_SIZE = 8;
const uint8_t ENC_MARKER = 0xff;

inline void encode_bytes(Bytes& out, Slice key) {
  size_t len = key.n, idx = 0;
  while (idx <= len) {
    size_t remain = len - idx, pad = 0;
    if (remain >= ENC_GROUP_SIZE) {
      out.insert(out.end(), key.p + idx, key.p + idx + ENC_GROUP_SIZE);
    } else {
      pad = ENC_GROUP_SIZE - remain;
      out.insert(out.end(), key.p + idx, key.p + len);
      out.insert(out.end(), pad, 0);
    }
    out.push_back((uint8_t)(ENC_MARKER - pad));
    idx += ENC_GROUP_SIZE;
  }
}
// returns consumed length, (size_t)-1 on error (unexpected eof / bad padding)
inline size_t decode_bytes(Slice data, Bytes* key) {
  size_t offset = 0;
  key->clear();
  for (;;) {
    size_t next = offset + ENC_GROUP_SIZE + 1;
    if (next > data.n) return (size_t)-1;
    const uint8_t* chunk = data.p + offset;
    offset = next;
    uint8_t marker = chunk[ENC_GROUP_SIZE];
    size_t pad = (size_t)(ENC_MARKER - marker);
    if (pad == 0) { key->insert(key->end(), chunk, chunk + ENC_GROUP_SIZE); continue; }
    if (pad > ENC_GROUP_SIZE) return (size_t)-1;
    key->insert(key->end(), chunk, chunk + (ENC_GROUP_SIZE - pad));
    for (size_t i = ENC_GROUP_SIZE - pad; i < ENC_GROUP_SIZE; ++i)
      if (chunk[i] != 0) return (size_t)-1;
    return offset;
  }
}
// bytes.rs get_first_encoded_len (asc): scan 9-byte groups until marker != 0xff
inline size_t memcmp_first_encoded_len(Slice s) {
  size_t idx = ENC_GROUP_SIZE;
  for (;;) {
    if (s.n < idx + 1) return s.n;
    if (s[idx] != ENC_MARKER) return idx + 1;
    idx += ENC_GROUP_SIZE + 1;
  }
}

// ---- Key: components/txn_types/src/types.rs:83-267 ----
inline Bytes key_from_raw(Slice raw) { Bytes b; encode_bytes(b, raw); return b; }             // :89-95
inline Bytes key_append_ts(Bytes k, uint64_t ts) { encode_u64_desc(k, ts); return k; }        // :152-155
inline bool key_split_ts(Slice k, Slice* user, uint64_t* ts) {                                // :211-228
  if (k.n < 8) return false;
  *user = Slice(k.p, k.n - 8);
  *ts = decode_u64_desc(k.p + k.n - 8);
  return true;
}
// types.rs:249-267 is_user_key_eq: ts_encoded_key minus 8-byte ts equals user_key
inline bool is_user_key_eq(Slice ts_encoded_key, Slice user_key) {
  if (ts_encoded_key.n != user_key.n + 8) return false;
  return memcmp(ts_encoded_key.p, user_key.p, user_key.n) == 0;
}

// ---- table record keys: components/tidb_query_datatype/src/codec/table.rs:26-34,187-218 ----
const size_t TBL_PREFIX_LEN = 11;  // 't' + i64 + "_r"
const size_t RECORD_ROW_KEY_LEN = 19;
inline Bytes encode_row_key(int64_t table_id, int64_t handle) {
  Bytes k;
  k.push_back('t');
  encode_i64(k, table_id);
  k.push_back('_'); k.push_back('r');
  encode_i64(k, handle);
  return k;
}
// table.rs:109-140 check_record_key; returns empty string on success else message
inline const char* check_record_key(Slice key) {
  if (key.n < 1) return "unexpected eof";
  if (key[0] != 't') return "record or index key expected";
  if (key.n < 9) return "unexpected eof";
  if (key.n < 11) return "unexpected eof";
  if (key[9] != '_' || key[10] != 'r') return "expected key sep type _r";
  return nullptr;
}
inline const char* decode_int_handle(Slice key, int64_t* handle) {  // table.rs:214-218
  const char* e = check_record_key(key);
  if (e) return e;
  if (key.n < TBL_PREFIX_LEN + 8) return "unexpected eof";
  *handle = decode_i64(key.p + TBL_PREFIX_LEN);
  return nullptr;
}
inline const char* decode_table_id(Slice key, int64_t* tid) {
  if (key.n < 1 || key[0] != 't') return "record key or index key expected";
  if (key.n < 9) return "unexpected eof";
  *tid = decode_i64(key.p + 1);
  return nullptr;
}

// ---- CRC-64/XZ (crc64fast 0.1.0, Cargo.lock:1695-1698; not vendored).  Published parameters:
// reflected poly 0xC96C5795D7870F42, init ~0, xorout ~0, check("123456789") = 0x995DC9BBDF1939FA.
struct Crc64Table {
  uint64_t t[256];
  Crc64Table() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint64_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0xC96C5795D7870F42ull : (c >> 1);
      t[i] = c;
    }
  }
};
inline const Crc64Table& crc64_table() { static Crc64Table t; return t; }
struct Crc64Digest {  // crc64fast::Digest: new(), write(), sum64()
  uint64_t state = ~0ull;
  void write(const uint8_t* p, size_t n) {
    const uint64_t* t = crc64_table().t;
    uint64_t c = state;
    for (size_t i = 0; i < n; ++i) c = t[(uint8_t)(c ^ p[i])] ^ (c >> 8);
    state = c;
  }
  uint64_t sum64() const { return ~state; }
};

}  // namespace orc
