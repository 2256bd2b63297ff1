// CPU ORACLE (test infrastructure, never on the product path): response encoding of the DAG runner.
//   encode_result_to_chunk          components/tidb_query_executors/src/runner.rs:1051-1088
//   LazyBatchColumnVec::encode      tidb_query_datatype/src/codec/batch/lazy_column_vec.rs:172-187  (TypeDefault: datum rows)
//   LazyBatchColumn::encode         codec/batch/lazy_column.rs:242-257 (Raw -> stored bytes, Decoded -> VectorValue::encode)
//   VectorValue::encode             codec/data_type/vector.rs:362-470, datum_codec.rs:248-287,330-350
//   DecimalEncoder::write_decimal   codec/mysql/decimal.rs:2025-2132, prec_and_frac :1043-1051
//   write_chunk_column              codec/chunk/column.rs:1052-1072 (TypeChunk), Column::new :50-70, append_* :446-496,640-644,840-
#pragma once
#include "orc_codec.h"
#include "orc_decimal.h"
#include "orc_exec.h"

namespace orc {

static const uint8_t DEC_DIG_2_BYTES[10] = {0, 1, 1, 2, 2, 3, 3, 4, 4, 4};

inline void dec_prec_and_frac(const Decimal& d, uint8_t* prec, uint8_t* frac) {  // decimal.rs:1043-1051
  int widx; uint8_t int_cnt;
  dec_remove_leading_zeroes(d, d.int_cnt, &widx, &int_cnt);
  uint8_t p = int_cnt + d.frac_cnt;
  *prec = p == 0 ? 1 : p; *frac = d.frac_cnt;
}

// decimal.rs:2025-2132.  `written` drives the sign-bit flip of the first byte (write_u8! / write_word! :1986-2020).
inline void dec_write(Bytes& out, const Decimal& d, uint8_t prec, uint8_t frac) {
  out.push_back(prec); out.push_back(frac);
  size_t written = 0;
  auto w_u8 = [&](uint8_t b) { if (written == 0) b ^= 0x80; out.push_back(b); written += 1; };
  auto w_word = [&](uint32_t word, int size) {
    uint8_t data[4] = {0, 0, 0, 0};
    for (int i = 0; i < size; ++i) data[i] = (uint8_t)(word >> (8 * (size - 1 - i)));
    if (written == 0) data[0] ^= 0x80;
    out.insert(out.end(), data, data + size);
    written += size;
  };
  uint32_t mask = d.negative ? 0xffffffffu : 0;
  int int_cnt = prec - frac;
  int int_word_cnt = int_cnt / 9, leading_digits = int_cnt - int_word_cnt * 9;
  int frac_word_cnt = frac / 9, trailing_digits = frac - frac_word_cnt * 9;
  int src_frac_word_cnt = d.frac_cnt / 9, src_trailing_digits = d.frac_cnt - src_frac_word_cnt * 9;
  int int_size = int_word_cnt * 4 + DEC_DIG_2_BYTES[leading_digits];
  int frac_size = frac_word_cnt * 4 + DEC_DIG_2_BYTES[trailing_digits];
  int src_frac_size = src_frac_word_cnt * 4 + DEC_DIG_2_BYTES[src_trailing_digits];
  int src_word_start_idx; uint8_t src_int_cnt_u8;
  dec_remove_leading_zeroes(d, d.int_cnt, &src_word_start_idx, &src_int_cnt_u8);
  int src_int_cnt = src_int_cnt_u8;
  if (src_int_cnt + src_frac_size == 0) { mask = 0; int_cnt = 1; }
  int src_int_word_cnt = src_int_cnt / 9, src_leading_digits = src_int_cnt - src_int_word_cnt * 9;
  int src_int_size = src_int_word_cnt * 4 + DEC_DIG_2_BYTES[src_leading_digits];
  if (int_cnt < src_int_cnt) {  // overflow: keep the low digits
    src_word_start_idx += src_int_word_cnt - int_word_cnt;
    if (src_leading_digits > 0) src_word_start_idx += 1;
    if (leading_digits > 0) src_word_start_idx -= 1;
    src_int_word_cnt = int_word_cnt;
    src_leading_digits = leading_digits;
  } else if (int_size > src_int_size) {
    for (int i = src_int_size; i < int_size; ++i) w_u8((uint8_t)mask);
  }
  if (frac_size < src_frac_size) {
    src_frac_word_cnt = frac_word_cnt;
    src_trailing_digits = trailing_digits;
  } else if (frac_size > src_frac_size && src_trailing_digits > 0) {
    if (frac_word_cnt == src_frac_word_cnt) { src_trailing_digits = trailing_digits; frac_size = src_frac_size; }
    else { src_frac_word_cnt += 1; src_trailing_digits = 0; }
  }
  if (src_leading_digits > 0) {
    int i = DEC_DIG_2_BYTES[src_leading_digits];
    uint32_t x = (d.word_buf[src_word_start_idx] % DEC_TEN_POW[src_leading_digits]) ^ mask;
    src_word_start_idx += 1;
    w_word(x, i);
  }
  int stop = src_word_start_idx + src_int_word_cnt + src_frac_word_cnt;
  while (src_word_start_idx < stop) { w_word(d.word_buf[src_word_start_idx] ^ mask, 4); src_word_start_idx += 1; }
  if (src_trailing_digits > 0) {
    int i = DEC_DIG_2_BYTES[src_trailing_digits];
    int lim = src_frac_word_cnt < frac_word_cnt ? 9 : trailing_digits;
    while (src_trailing_digits < lim && DEC_DIG_2_BYTES[src_trailing_digits] == i) src_trailing_digits += 1;
    uint32_t x = (d.word_buf[src_word_start_idx] / DEC_TEN_POW[9 - src_trailing_digits]) ^ mask;
    w_word(x, i);
  }
  if (frac_size > src_frac_size) {
    for (int k = src_frac_size; k < frac_size && written < (size_t)(int_size + frac_size); ++k) w_u8((uint8_t)mask);
  }
}

// datum_codec.rs:248-287
inline void write_datum_null(Bytes& o) { o.push_back(0); }
inline void write_datum_i64(Bytes& o, int64_t v) { o.push_back(3); put_u64_be(o, (uint64_t)v ^ SIGN_MARK); }
inline void write_datum_u64(Bytes& o, uint64_t v) { o.push_back(4); put_u64_be(o, v); }
inline void write_datum_f64(Bytes& o, double v) { o.push_back(5); put_u64_be(o, encode_f64_to_cmp_u64(v)); }
inline void write_datum_decimal(Bytes& o, const Decimal& d) {
  o.push_back(6);
  uint8_t prec, frac;
  dec_prec_and_frac(d, &prec, &frac);
  dec_write(o, d, prec, frac);
}

// One cell of a batch column into the TypeDefault row stream (lazy_column.rs:242-257 + vector.rs:362-470).
inline void encode_cell_default(Bytes& o, const LazyColumn& c, size_t r, const FieldType& ft, const std::vector<Decimal>* dec_cells) {
  if (!c.decoded) { Slice s = c.raw_get(r); o.insert(o.end(), s.p, s.p + s.n); return; }
  if (!c.nn[r]) { write_datum_null(o); return; }
  if (c.et == ET_TIME) {  // a decoded DateTime goes out as its packed u64 (Time::to_packed_u64, datum.rs write_datum)
    const uint64_t b = (uint64_t)c.i64[r];
    const uint64_t year = (b >> 50) & 0x3fff, month = (b >> 46) & 15, day = (b >> 41) & 31, hour = (b >> 36) & 31, minute = (b >> 30) & 63, second = (b >> 24) & 63, micro = (b >> 4) & 0xfffff;
    write_datum_u64(o, (((((year * 13 + month) << 5) | day) << 17) | (hour << 12) | (minute << 6) | second) << 24 | micro);
  } else if (c.et == ET_DURATION) { o.push_back(DURATION_FLAG); uint64_t u = (uint64_t)c.i64[r] ^ 0x8000000000000000ull; for (int k = 7; k >= 0; --k) o.push_back((uint8_t)(u >> (8 * k))); }
  else if (c.et == ET_REAL) write_datum_f64(o, c.f64[r]);
  else if (c.et == ET_DECIMAL) write_datum_decimal(o, (*dec_cells)[(size_t)c.i64[r]]);
  else if (ft.is_unsigned()) write_datum_u64(o, (uint64_t)c.i64[r]);
  else write_datum_i64(o, c.i64[r]);
}

// TypeChunk block of one fully decoded result column (column.rs:50-70 fixed lengths, :446-496 appends, :1052-1072 layout):
// u32 length | u32 null_cnt | bitmap (only when null_cnt > 0; bit = 1 -> non-null) | fixed-width cells (NULL cells zero)
inline void encode_column_chunk(Bytes& o, int tp, const std::vector<uint8_t>& nn, const int64_t* i64, const double* f64, const Decimal* dec) {
  size_t n = nn.size(), null_cnt = 0;
  for (uint8_t b : nn) null_cnt += !b;
  auto le32 = [&](uint32_t v) { for (int i = 0; i < 4; ++i) o.push_back((uint8_t)(v >> (8 * i))); };
  le32((uint32_t)n); le32((uint32_t)null_cnt);
  if (null_cnt > 0) {
    size_t nb = (n + 7) / 8, base = o.size();
    o.resize(base + nb, 0);
    for (size_t i = 0; i < n; ++i) if (nn[i]) o[base + (i >> 3)] |= (uint8_t)(1u << (i & 7));
  }
  for (size_t i = 0; i < n; ++i) {
    if (dec) { const uint8_t* p = (const uint8_t*)&dec[i]; Decimal z; memset(&z, 0, sizeof(z)); if (!nn[i]) p = (const uint8_t*)&z; o.insert(o.end(), p, p + 40); }
    else if (tp == B2_TP_FLOAT) { float f = nn[i] ? (float)f64[i] : 0.0f; uint32_t u; memcpy(&u, &f, 4); le32(u); }
    else { uint64_t u = 0; if (nn[i]) { if (f64) memcpy(&u, &f64[i], 8); else u = (uint64_t)i64[i]; } for (int k = 0; k < 8; ++k) o.push_back((uint8_t)(u >> (8 * k))); }
  }
}

}  // namespace orc
