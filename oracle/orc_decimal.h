// TEST INFRASTRUCTURE — CPU oracle, not product code.
// MySQL Decimal restatement: components/tidb_query_datatype/src/codec/mysql/decimal.rs
//   struct :927-942, From<i64>/<u64> :1787-1815, do_add :492-589, calc_sub_carry/do_sub :265-448,
//   Add :2340-2353, to_string_value :1925-1969, Ord :2323-2338, is_zero :1743-1746.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>

namespace orc {

const uint8_t DEC_WORD_BUF_LEN = 9;
const uint8_t DEC_DIGITS_PER_WORD = 9;
const uint32_t DEC_WORD_BASE = 1000000000u;
const uint32_t DEC_WORD_MAX = DEC_WORD_BASE - 1;
const uint32_t DEC_DIG_MASK = 100000000u;
static const uint32_t DEC_TEN_POW[10] = {1, 10, 100, 1000, 10000, 100000, 1000000, 10000000, 100000000, 1000000000};

enum DecRes { DEC_OK = 0, DEC_TRUNCATED = 1, DEC_OVERFLOW = 2 };

struct Decimal {  // #[repr(C)], 40 bytes
  uint8_t int_cnt, frac_cnt, result_frac_cnt, negative;
  uint32_t word_buf[9];
};
static_assert(sizeof(Decimal) == 40, "DECIMAL_STRUCT_SIZE");

inline uint8_t dec_word_cnt(int len) {  // word_cnt! macro :147-166
  if (len > 0 && len > DEC_DIGITS_PER_WORD * DEC_WORD_BUF_LEN) return DEC_WORD_BUF_LEN + 1;
  if (len <= 0) return 0;
  return (uint8_t)((len + DEC_DIGITS_PER_WORD - 1) / DEC_DIGITS_PER_WORD);
}
inline Decimal dec_new(uint8_t int_cnt, uint8_t frac_cnt, bool negative) {
  Decimal d;
  memset(&d, 0, sizeof(d));
  d.int_cnt = int_cnt; d.frac_cnt = frac_cnt; d.result_frac_cnt = frac_cnt; d.negative = negative;
  return d;
}
inline Decimal dec_zero() { return dec_new(1, 0, false); }

inline Decimal dec_from_u64(uint64_t u) {  // :1799-1815
  uint64_t x = u;
  uint8_t word_idx = 1;
  while (x >= DEC_WORD_BASE) { word_idx++; x /= DEC_WORD_BASE; }
  Decimal d = dec_new(word_idx * DEC_DIGITS_PER_WORD, 0, false);
  x = u;
  while (word_idx > 0) {
    word_idx--;
    d.word_buf[word_idx] = (uint32_t)(x % DEC_WORD_BASE);
    x /= DEC_WORD_BASE;
  }
  return d;
}
inline Decimal dec_from_i64(int64_t i) {  // :1787-1797
  if (i < 0) {
    Decimal d = dec_from_u64((uint64_t)0 - (uint64_t)i);
    d.negative = 1;
    return d;
  }
  return dec_from_u64((uint64_t)i);
}

inline bool dec_is_zero(const Decimal& d) {
  int len = dec_word_cnt(d.int_cnt) + dec_word_cnt(d.frac_cnt);
  for (int i = 0; i < len; ++i) if (d.word_buf[i]) return false;
  return true;
}

inline void dec_add_word(uint32_t a, uint32_t b, uint32_t* carry, uint32_t* res) {  // fn add :215-225
  uint32_t sum = a + b + *carry;
  if (sum >= DEC_WORD_BASE) { *res = sum - DEC_WORD_BASE; *carry = 1; } else { *res = sum; *carry = 0; }
}
inline void dec_sub_word(uint32_t l, uint32_t r, int32_t* carry, uint32_t* res) {  // fn sub :239-249
  int32_t diff = (int32_t)l - (int32_t)r - *carry;
  if (diff < 0) { *carry = 1; *res = (uint32_t)(diff + (int32_t)DEC_WORD_BASE); } else { *carry = 0; *res = (uint32_t)diff; }
}
// fix_word_cnt_err :228-236
inline DecRes dec_fix_word_cnt(uint8_t* int_wc, uint8_t* frac_wc, uint8_t buf_len) {
  if (*int_wc + *frac_wc > buf_len) {
    if (*int_wc > buf_len) { *int_wc = buf_len; *frac_wc = 0; return DEC_OVERFLOW; }
    *frac_wc = buf_len - *int_wc;
    return DEC_TRUNCATED;
  }
  return DEC_OK;
}
inline Decimal dec_max(uint8_t prec, uint8_t frac_cnt) {  // max_decimal :450-478 (frac part elided: only frac 0 used)
  uint8_t int_cnt = prec - frac_cnt;
  Decimal res = dec_new(int_cnt, frac_cnt, false);
  int idx = 0;
  if (int_cnt > 0) {
    uint8_t first = int_cnt % DEC_DIGITS_PER_WORD;
    if (first > 0) res.word_buf[idx++] = DEC_TEN_POW[first] - 1;
    for (int i = 0; i < int_cnt / DEC_DIGITS_PER_WORD; ++i) res.word_buf[idx++] = DEC_WORD_MAX;
  }
  for (int i = 0; i < frac_cnt / DEC_DIGITS_PER_WORD; ++i) res.word_buf[idx++] = DEC_WORD_MAX;
  return res;
}

// do_add :492-589 (lhs.negative == rhs.negative)
inline DecRes dec_do_add(const Decimal* lhs, const Decimal* rhs, Decimal* out) {
  uint8_t l_int = dec_word_cnt(lhs->int_cnt), l_frac = dec_word_cnt(lhs->frac_cnt);
  uint8_t r_int = dec_word_cnt(rhs->int_cnt), r_frac = dec_word_cnt(rhs->frac_cnt);
  uint8_t int_to = std::max(l_int, r_int), frac_to = std::max(l_frac, r_frac);
  uint32_t x = l_int > r_int ? lhs->word_buf[0] : (l_int < r_int ? rhs->word_buf[0] : lhs->word_buf[0] + rhs->word_buf[0]);
  if (x > DEC_WORD_MAX - 1) int_to += 1;
  DecRes st = dec_fix_word_cnt(&int_to, &frac_to, DEC_WORD_BUF_LEN);
  if (st == DEC_OVERFLOW) { *out = dec_max(DEC_WORD_BUF_LEN * DEC_DIGITS_PER_WORD, 0); return DEC_OVERFLOW; }
  int idx_to = int_to + frac_to;
  Decimal res = dec_new(int_to * DEC_DIGITS_PER_WORD, std::max(lhs->frac_cnt, rhs->frac_cnt), lhs->negative);
  res.word_buf[0] = 0;
  if (st != DEC_OK) {
    res.frac_cnt = std::min<uint8_t>(frac_to * DEC_DIGITS_PER_WORD, res.frac_cnt);
    l_frac = std::min(frac_to, l_frac); r_frac = std::min(r_frac, frac_to);
    l_int = std::min(l_int, int_to); r_int = std::min(r_int, int_to);
  }
  int l_idx, r_idx, l_stop, r_stop;
  bool exchanged;
  if (l_frac > r_frac) {
    l_idx = l_int + l_frac; l_stop = l_int + r_frac; r_idx = r_int + r_frac;
    r_stop = l_int > r_int ? l_int - r_int : 0;
    exchanged = false;
  } else {
    l_idx = r_int + r_frac; l_stop = r_int + l_frac; r_idx = l_int + l_frac;
    r_stop = r_int > l_int ? r_int - l_int : 0;
    std::swap(lhs, rhs);
    exchanged = true;
  }
  while (l_idx > l_stop) { idx_to--; l_idx--; res.word_buf[idx_to] = lhs->word_buf[l_idx]; }
  uint32_t carry = 0;
  while (l_idx > r_stop) {
    l_idx--; r_idx--; idx_to--;
    dec_add_word(lhs->word_buf[l_idx], rhs->word_buf[r_idx], &carry, &res.word_buf[idx_to]);
  }
  l_stop = 0;
  if (l_int > r_int) {
    l_idx = l_int - r_int;
    if (exchanged) std::swap(lhs, rhs);
  } else {
    l_idx = r_int - l_int;
    if (!exchanged) std::swap(lhs, rhs);
  }
  while (l_idx > l_stop) { idx_to--; l_idx--; dec_add_word(lhs->word_buf[l_idx], 0, &carry, &res.word_buf[idx_to]); }
  if (carry > 0) { idx_to--; res.word_buf[idx_to] = 1; }
  *out = res;
  return st;
}

struct DecSubTmp { int start, int_wc, frac_wc; };
// calc_sub_carry :265-342. carry: -1 = None (equal), 0 = |l|>|r|, 1 = |l|<|r|
inline int dec_calc_sub_carry(const Decimal* lhs, const Decimal* rhs, uint8_t* frac_word_to, DecSubTmp* l, DecSubTmp* r) {
  int l_int = dec_word_cnt(lhs->int_cnt), l_frac = dec_word_cnt(lhs->frac_cnt);
  int r_int = dec_word_cnt(rhs->int_cnt), r_frac = dec_word_cnt(rhs->frac_cnt);
  *frac_word_to = (uint8_t)std::max(l_frac, r_frac);
  int l_stop = l_int, l_idx = 0;
  while (l_idx < l_stop && lhs->word_buf[l_idx] == 0) l_idx++;
  int l_start = l_idx; l_int = l_stop - l_idx;
  int r_stop = r_int, r_idx = 0;
  while (r_idx < r_stop && rhs->word_buf[r_idx] == 0) r_idx++;
  int r_start = r_idx; r_int = r_stop - r_idx;
  int carry;
  if (r_int > l_int) carry = 1;
  else if (r_int < l_int) carry = 0;
  else {
    int l_end = l_stop + l_frac - 1, r_end = r_stop + r_frac - 1;
    while (l_idx <= l_end && lhs->word_buf[l_end] == 0) l_end--;
    while (r_idx <= r_end && rhs->word_buf[r_end] == 0) r_end--;
    l_frac = std::max(0, l_end + 1 - l_stop);
    r_frac = std::max(0, r_end + 1 - r_stop);
    while (l_idx <= l_end && r_idx <= r_end && lhs->word_buf[l_idx] == rhs->word_buf[r_idx]) { l_idx++; r_idx++; }
    if (l_idx <= l_end) {
      carry = (r_idx <= r_end && rhs->word_buf[r_idx] > lhs->word_buf[l_idx]) ? 1 : 0;
    } else if (r_idx <= r_end) carry = 1;
    else carry = -1;
  }
  l->start = l_start; l->int_wc = l_int; l->frac_wc = l_frac;
  r->start = r_start; r->int_wc = r_int; r->frac_wc = r_frac;
  return carry;
}

// do_sub :345-448: |lhs| - |rhs| with sign handling, used when signs differ in Add
inline DecRes dec_do_sub(const Decimal* lhs, const Decimal* rhs, Decimal* out) {
  uint8_t frac_word_to;
  DecSubTmp lt, rt;
  int carry0 = dec_calc_sub_carry(lhs, rhs, &frac_word_to, &lt, &rt);
  if (carry0 < 0) { *out = dec_zero(); return DEC_OK; }
  bool negative;
  if (carry0 > 0) { std::swap(lhs, rhs); std::swap(lt, rt); negative = !rhs->negative; } else negative = lhs->negative;
  uint8_t l_int_wc = (uint8_t)lt.int_wc;
  DecRes st = dec_fix_word_cnt(&l_int_wc, &frac_word_to, DEC_WORD_BUF_LEN);
  lt.int_wc = l_int_wc;
  int idx_to = lt.int_wc + frac_word_to;
  uint8_t frac_cnt = std::max(lhs->frac_cnt, rhs->frac_cnt);
  uint8_t int_cnt = (uint8_t)(lt.int_wc * DEC_DIGITS_PER_WORD);
  if (st != DEC_OK) {
    frac_cnt = std::min<uint8_t>(frac_cnt, frac_word_to * DEC_DIGITS_PER_WORD);
    lt.frac_wc = std::min<int>(lt.frac_wc, frac_word_to);
    rt.frac_wc = std::min<int>(rt.frac_wc, frac_word_to);
    rt.int_wc = std::min(rt.int_wc, lt.int_wc);
  }
  int32_t carry = 0;
  Decimal res = dec_new(int_cnt, frac_cnt, negative);
  int l_idx = lt.start + lt.int_wc + lt.frac_wc;
  int r_idx = rt.start + rt.int_wc + rt.frac_wc;
  if (lt.frac_wc > rt.frac_wc) {
    int l_stop = lt.start + lt.int_wc + rt.frac_wc;
    if (lt.frac_wc < frac_word_to) idx_to -= frac_word_to - lt.frac_wc;
    while (l_idx > l_stop) { idx_to--; l_idx--; res.word_buf[idx_to] = lhs->word_buf[l_idx]; }
  } else {
    int r_stop = rt.start + rt.int_wc + lt.frac_wc;
    if (frac_word_to > rt.frac_wc) idx_to -= frac_word_to - rt.frac_wc;
    while (r_idx > r_stop) { idx_to--; r_idx--; dec_sub_word(0, rhs->word_buf[r_idx], &carry, &res.word_buf[idx_to]); }
  }
  while (r_idx > rt.start) {
    idx_to--; l_idx--; r_idx--;
    dec_sub_word(lhs->word_buf[l_idx], rhs->word_buf[r_idx], &carry, &res.word_buf[idx_to]);
  }
  while (carry > 0 && l_idx > lt.start) { idx_to--; l_idx--; dec_sub_word(lhs->word_buf[l_idx], 0, &carry, &res.word_buf[idx_to]); }
  while (l_idx > lt.start) { idx_to--; l_idx--; res.word_buf[idx_to] = lhs->word_buf[l_idx]; }
  *out = res;
  return st;
}

// impl Add :2340-2353
inline DecRes dec_add(const Decimal& a, const Decimal& b, Decimal* out) {
  uint8_t rfc = std::max(a.result_frac_cnt, b.result_frac_cnt);
  DecRes st = (a.negative == b.negative) ? dec_do_add(&a, &b, out) : dec_do_sub(&a, &b, out);
  out->result_frac_cnt = rfc;
  return st;
}

inline int dec_cmp(const Decimal& a, const Decimal& b) {  // Ord :2323-2338
  if (a.negative == b.negative) {
    uint8_t f; DecSubTmp l, r;
    int carry = dec_calc_sub_carry(&a, &b, &f, &l, &r);
    if (carry < 0) return 0;
    return ((carry > 0) == (bool)a.negative) ? 1 : -1;
  }
  return a.negative ? -1 : 1;
}

// remove_leading_zeroes :1003-1018
inline void dec_remove_leading_zeroes(const Decimal& d, uint8_t prec, int* word_idx, uint8_t* cnt_out) {
  int cnt = prec;
  int i = ((cnt + DEC_DIGITS_PER_WORD - 1) % DEC_DIGITS_PER_WORD) + 1;
  int widx = 0;
  while (cnt > 0 && d.word_buf[widx] == 0) { cnt -= i; i = DEC_DIGITS_PER_WORD; widx++; }
  if (cnt > 0) {
    int k = (cnt - 1) % DEC_DIGITS_PER_WORD, c = 0;
    while (DEC_TEN_POW[k] > d.word_buf[widx]) { k--; c++; }  // count_leading_zeroes :200-207
    cnt -= c;
  }
  *word_idx = widx; *cnt_out = (uint8_t)cnt;
}

inline std::string dec_to_string(const Decimal& d) {  // to_string_value :1925-1969
  uint8_t frac_cnt = d.frac_cnt, int_cnt;
  int word_start;
  dec_remove_leading_zeroes(d, d.int_cnt, &word_start, &int_cnt);
  if (int_cnt + frac_cnt == 0) { int_cnt = 1; word_start = 0; }
  std::string buf;
  if (d.negative) buf.push_back('-');
  if (int_cnt > 0) {
    size_t base = buf.size();
    size_t idx = base + int_cnt;
    int widx = word_start + dec_word_cnt(int_cnt);
    buf.resize(idx, '0');
    while (idx > base) {
      widx--;
      uint32_t x = d.word_buf[widx];
      int n = (int)std::min<size_t>(idx - base, DEC_DIGITS_PER_WORD);
      for (int k = 0; k < n; ++k) { idx--; buf[idx] = (char)('0' + x % 10); x /= 10; }
    }
  } else buf.push_back('0');
  if (frac_cnt > 0) {
    buf.push_back('.');
    int widx = word_start + dec_word_cnt(int_cnt);
    size_t exp_idx = buf.size() + frac_cnt;
    while (buf.size() < exp_idx) {
      uint32_t x = d.word_buf[widx];
      int n = (int)std::min<size_t>(exp_idx - buf.size(), DEC_DIGITS_PER_WORD);
      for (int k = 0; k < n; ++k) { buf.push_back((char)('0' + x / DEC_DIG_MASK)); x = (x % DEC_DIG_MASK) * 10; }
      widx++;
    }
  }
  return buf;
}

}  // namespace orc
