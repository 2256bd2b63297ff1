// CPU ORACLE (test infrastructure, never on the product path): raw datum -> chunk column cells for the column types the
// executors never decode (they stay LazyBatchColumn::Raw until the response is encoded).
//   Column::from_raw_datums            tidb_query_datatype/src/codec/chunk/column.rs:72-151
//   append_bytes_datum                 :697-724     append_time_datum      :742-778
//   append_duration_datum              :797-827     append_decimal_datum   :846-870     append_json_datum :890-913
//   DecimalDecoder::read_decimal       codec/mysql/decimal.rs:2204-2289 (read_word :2159-2200)
//   Time::from_packed_u64              codec/mysql/time/mod.rs:2002-2043, bit field :167-196, set_tt :1989-2000, set_fsp :1939-1947
//   write_chunk_column                 chunk/column.rs:1052-1072 (var-length columns: (n + 1) i64 offsets before the data)
#pragma once
#include "orc_codec.h"
#include "orc_decimal.h"
#include "orc_exec.h"

namespace orc {

enum RawKind { RK_NONE = 0, RK_BYTES, RK_TIME, RK_DURATION, RK_DECIMAL, RK_JSON };
inline RawKind raw_kind_of(int tp) {  // EvalType::try_from(FieldTypeTp), def/eval_type.rs:53-95
  switch (tp) {
    case B2_TP_VARCHAR: case B2_TP_VARSTRING: case B2_TP_STRING: case B2_TP_BLOB: case 0xf9: case 0xfa: case 0xfb: case 0xff: return RK_BYTES;
    case B2_TP_DATE: case B2_TP_DATETIME: case B2_TP_TIMESTAMP: return RK_TIME;
    case B2_TP_DURATION: return RK_DURATION;
    case B2_TP_NEWDECIMAL: return RK_DECIMAL;
    case B2_TP_JSON: return RK_JSON;
    default: return RK_NONE;
  }
}

struct RawChunkCol {
  RawKind kind = RK_NONE;
  std::vector<uint8_t> nn;
  Bytes data;                         // fixed cells (8 or 40 bytes each, NULL cells zero) or the byte heap
  std::vector<int64_t> offsets{0};    // var-length kinds
  bool var() const { return kind == RK_BYTES || kind == RK_JSON; }
  size_t fixed_len() const { return kind == RK_DECIMAL ? 40 : 8; }
  void clear() { nn.clear(); data.clear(); offsets.assign(1, 0); }
};

inline bool dec_read(Slice s, Decimal* out, std::string* err) {  // decimal.rs:2204-2289
  if (s.n < 3) { *err = "decimal too short"; return false; }
  uint8_t prec = s[0], frac_cnt = s[1];
  if (prec < frac_cnt) { *err = "invalid decimal"; return false; }
  s = s.sub(2);
  static const uint8_t D2B[10] = {0, 1, 1, 2, 2, 3, 3, 4, 4, 4};
  uint8_t int_cnt = prec - frac_cnt;
  int int_word_cnt = int_cnt / 9, leading = int_cnt - int_word_cnt * 9, frac_word_cnt = frac_cnt / 9, trailing = frac_cnt - frac_word_cnt * 9;
  int int_word_to = int_word_cnt + (leading > 0), frac_word_to = frac_word_cnt + (trailing > 0);
  uint32_t mask = (s[0] & 0x80) ? 0 : 0xffffffffu;
  if (int_word_to + frac_word_to > 9) { *err = "decoding decimal failed"; return false; }
  Decimal d = dec_new(int_cnt, frac_cnt, mask != 0);
  d.result_frac_cnt = frac_cnt;
  bool is_first = true, ok = true;
  auto read_word = [&](int size) -> uint32_t {  // :2159-2200
    if ((int)s.n < size) { ok = false; return 0; }
    uint8_t first = s[0];
    if (is_first) { first ^= 0x80; is_first = false; }
    uint32_t r;
    switch (size) {
      case 1: r = (uint32_t)(int32_t)(int8_t)first; break;
      case 2: r = (uint32_t)(((int32_t)(int8_t)first << 8) + (int32_t)s[1]); break;
      case 3: r = (first & 128) ? ((255u << 24) | ((uint32_t)first << 16) | ((uint32_t)s[1] << 8) | s[2]) : (((uint32_t)first << 16) | ((uint32_t)s[1] << 8) | s[2]); break;
      default: r = (uint32_t)(((int32_t)(int8_t)first << 24) + ((int32_t)s[1] << 16) + ((int32_t)s[2] << 8) + (int32_t)s[3]); break;
    }
    s = s.sub(size);
    return r;
  };
  int w = 0;
  if (leading > 0) {
    d.word_buf[w] = read_word(D2B[leading]) ^ mask;
    if (!ok) { *err = "unexpected eof"; return false; }
    if (d.word_buf[w] >= DEC_TEN_POW[leading + 1]) { *err = "invalid leading digits for decimal number"; return false; }
    if (d.word_buf[w] != 0) w++; else d.int_cnt -= (uint8_t)leading;
  }
  for (int i = 0; i < int_word_cnt; ++i) {
    d.word_buf[w] = read_word(4) ^ mask;
    if (!ok) { *err = "unexpected eof"; return false; }
    if (d.word_buf[w] > 999999999u) { *err = "invalid int part for decimal number"; return false; }
    if (w > 0 || d.word_buf[w] != 0) w++; else d.int_cnt -= 9;
  }
  for (int i = 0; i < frac_word_cnt; ++i) {
    d.word_buf[w] = read_word(4) ^ mask;
    if (!ok) { *err = "unexpected eof"; return false; }
    if (d.word_buf[w] > 999999999u) { *err = "invalid frac part decimal number"; return false; }
    w++;
  }
  if (trailing > 0) {
    uint32_t x = read_word(D2B[trailing]) ^ mask;
    if (!ok) { *err = "unexpected eof"; return false; }
    uint64_t v = (uint64_t)x * DEC_TEN_POW[9 - trailing];
    if (v > 999999999ull) { *err = "invalid trailing digits for decimal number"; return false; }
    d.word_buf[w] = (uint32_t)v;
  }
  if (d.int_cnt == 0 && d.frac_cnt == 0) d = dec_zero();
  d.result_frac_cnt = frac_cnt;
  *out = d;
  return true;
}

inline void put_le64(Bytes& o, uint64_t v) { for (int k = 0; k < 8; ++k) o.push_back((uint8_t)(v >> (8 * k))); }

// one cell of Column::from_raw_datums
inline bool append_raw_datum(RawChunkCol& c, Slice d, const FieldType& ft, std::string* err) {
  if (d.empty()) { *err = "Failed to decode datum flag"; return false; }
  const uint8_t flag = d[0];
  Slice p = d.sub(1);
  auto null_cell = [&]() { c.nn.push_back(0); if (c.var()) c.offsets.push_back((int64_t)c.data.size()); else c.data.insert(c.data.end(), c.fixed_len(), 0); };
  if (flag == NIL_FLAG) { null_cell(); return true; }
  switch (c.kind) {
    case RK_BYTES:
      if (flag == COMPACT_BYTES_FLAG) {
        int64_t vn; size_t n = decode_var_i64(p, &vn);
        if (!n || vn < 0 || p.n - n < (size_t)vn) { *err = "unexpected eof"; return false; }
        c.data.insert(c.data.end(), p.p + n, p.p + n + vn);
      } else if (flag == BYTES_FLAG) {
        Bytes out; if (decode_bytes(p, &out) == (size_t)-1) { *err = "unexpected eof"; return false; }
        c.data.insert(c.data.end(), out.begin(), out.end());
      } else { *err = "Unsupported datum flag " + std::to_string(flag) + " for Bytes vector"; return false; }
      c.offsets.push_back((int64_t)c.data.size());
      break;
    case RK_JSON:
      if (flag != JSON_FLAG) { *err = "Unsupported datum flag " + std::to_string(flag) + " for Json vector"; return false; }
      c.data.insert(c.data.end(), p.p, p.p + p.n);  // write_json_to_chunk_by_datum_payload: the payload as it is
      c.offsets.push_back((int64_t)c.data.size());
      break;
    case RK_TIME: {
      uint64_t v;
      if (flag == UINT_FLAG) { if (p.n < 8) { *err = "unexpected eof"; return false; } v = get_u64_be(p.p); }
      else if (flag == VAR_UINT_FLAG) { if (!decode_var_u64(p, &v)) { *err = "unexpected eof"; return false; } }
      else { *err = "Unsupported datum flag " + std::to_string(flag) + " for DateTime vector."; return false; }
      uint64_t bits;
      if (!time_from_packed(v, ft.tp, ft.decimal, &bits, err)) return false;
      put_le64(c.data, bits);
      break;
    }
    case RK_DURATION: {
      int64_t v;
      if (flag == DURATION_FLAG) { if (p.n < 8) { *err = "unexpected eof"; return false; } v = decode_i64(p.p); }
      else if (flag == VAR_INT_FLAG) { if (!decode_var_i64(p, &v)) { *err = "unexpected eof"; return false; } }
      else { *err = "Unsupported datum flag " + std::to_string(flag) + " for Duration vector"; return false; }
      put_le64(c.data, (uint64_t)v);
      break;
    }
    case RK_DECIMAL: {
      if (flag != DECIMAL_FLAG) { *err = "Unsupported datum flag " + std::to_string(flag) + " for Decimal vector"; return false; }
      Decimal dv;
      if (!dec_read(p, &dv, err)) return false;
      const uint8_t* b = (const uint8_t*)&dv;
      c.data.insert(c.data.end(), b, b + 40);
      break;
    }
    default: *err = "unsupported column type"; return false;
  }
  c.nn.push_back(1);
  return true;
}

// write_chunk_column of one such column
inline void encode_raw_chunk(Bytes& o, const RawChunkCol& c) {
  size_t n = c.nn.size(), null_cnt = 0;
  for (uint8_t b : c.nn) null_cnt += !b;
  auto le32 = [&](uint32_t v) { for (int i = 0; i < 4; ++i) o.push_back((uint8_t)(v >> (8 * i))); };
  le32((uint32_t)n); le32((uint32_t)null_cnt);
  if (null_cnt > 0) {
    size_t nb = (n + 7) / 8, base = o.size();
    o.resize(base + nb, 0);
    for (size_t i = 0; i < n; ++i) if (c.nn[i]) o[base + (i >> 3)] |= (uint8_t)(1u << (i & 7));
  }
  if (c.var()) for (int64_t v : c.offsets) put_le64(o, (uint64_t)v);
  o.insert(o.end(), c.data.begin(), c.data.end());
}

}  // namespace orc
