// TEST INFRASTRUCTURE — host-side synthetic region generator (bench.py's CPU arm and the bench-scale parity check).
//
// Produces, on the host and without any CUDA call, the same bytes as the device generator of the product library
// (tikv_b200/csrc/kernels.cu gen_sizes / gen_write; spec = b2_gen_spec of include/b2_copr.h), so that the reference arm
// needs nothing from libb2copr.so and the oracle can be run on exactly the rows a GPU request covers.  Table layout as in
// record key 't' i64cmp(table_id) "_r" i64cmp(handle) (table.rs:187-193), memcomparable + !commit_ts
// (types.rs:152-155), write record `P varint(start_ts) v len row` (write.rs:363-393), row format v2 (row_slice.rs:74-115)
// or v1 (table_scan_executor.rs:200-247).  tests/test_gpu_parity.py checks byte equality with the device generator.
#pragma once
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/b2_copr.h"

namespace orc {

struct GenBlock {
  std::vector<uint8_t> keys, vals;
  std::vector<uint32_t> koff, voff;
  uint64_t n_user_keys = 0;
  b2_cf_block view() const { b2_cf_block b; b.keys = keys.data(); b.key_offs = koff.data(); b.vals = vals.data(); b.val_offs = voff.data(); b.n = (uint32_t)(koff.size() - 1); b._pad = 0; return b; }
};

inline uint64_t gen_mix64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }
inline uint64_t gen_mix(uint64_t seed, uint64_t handle, uint64_t salt) { return gen_mix64(seed ^ (handle * 0x9E3779B97F4A7C15ull) ^ ((salt + 1) * 0xBF58476D1CE4E5B9ull)); }

struct GenRowKind { int kind; uint32_t entries; };  // 0 plain, 1 extra versions, 2 delete, 3 lock record
inline GenRowKind gen_row_kind(const b2_gen_spec& s, uint64_t handle) {
  uint64_t r = gen_mix(s.seed, handle, 1000) % 1000000ull;
  if (r < s.extra_versions_per_million) return {1, 3};
  if (r < (uint64_t)s.extra_versions_per_million + s.delete_per_million) return {2, 2};
  if (r < (uint64_t)s.extra_versions_per_million + s.delete_per_million + s.lock_rec_per_million) return {3, 2};
  return {0, 1};
}
inline bool gen_value(const b2_gen_spec& s, uint64_t handle, uint32_t c, uint64_t version_salt, int64_t* v) {
  if (s.null_per_million && s.null_per_million[c] && gen_mix(s.seed, handle, 500 + c) % 1000000ull < s.null_per_million[c]) return false;
  uint64_t x = gen_mix(s.seed + version_salt, handle, c);
  uint64_t range = s.col_range ? s.col_range[c] : 0;
  int64_t lo = s.col_lo ? s.col_lo[c] : 0;
  *v = range ? (int64_t)((uint64_t)lo + x % range) : (int64_t)x;
  return true;
}
inline uint32_t gen_int_width(int64_t v) { return (v >= -128 && v <= 127) ? 1 : (v >= -32768 && v <= 32767) ? 2 : (v >= -2147483648ll && v <= 2147483647ll) ? 4 : 8; }
inline void gen_put_varint(std::vector<uint8_t>& o, uint64_t v) { while (v >= 0x80) { o.push_back((uint8_t)(v | 0x80)); v >>= 7; } o.push_back((uint8_t)v); }
inline uint64_t gen_zigzag(int64_t v) { uint64_t u = (uint64_t)v << 1; return v < 0 ? ~u : u; }

inline void gen_row_bytes(const b2_gen_spec& s, uint64_t handle, uint64_t version_salt, std::vector<uint8_t>& o) {
  if (s.row_format == 2) {
    std::vector<uint8_t> ids, nulls, offs, vals;
    for (uint32_t c = 0; c < s.n_cols; ++c) {
      int64_t v;
      if (gen_value(s, handle, c, version_salt, &v)) {
        uint32_t w = gen_int_width(v);
        ids.push_back((uint8_t)(c + 1));
        for (uint32_t i = 0; i < w; ++i) vals.push_back((uint8_t)((uint64_t)v >> (8 * i)));
        offs.push_back((uint8_t)vals.size()); offs.push_back((uint8_t)(vals.size() >> 8));
      } else nulls.push_back((uint8_t)(c + 1));
    }
    o.push_back(128); o.push_back(0);
    o.push_back((uint8_t)ids.size()); o.push_back((uint8_t)(ids.size() >> 8)); o.push_back((uint8_t)nulls.size()); o.push_back((uint8_t)(nulls.size() >> 8));
    o.insert(o.end(), ids.begin(), ids.end()); o.insert(o.end(), nulls.begin(), nulls.end());
    o.insert(o.end(), offs.begin(), offs.end()); o.insert(o.end(), vals.begin(), vals.end());
    return;
  }
  for (uint32_t c = 0; c < s.n_cols; ++c) {  // v1: [VAR_INT colid][VAR_INT zigzag | NIL]
    int64_t v;
    bool nonnull = gen_value(s, handle, c, version_salt, &v);
    o.push_back(8); gen_put_varint(o, gen_zigzag((int64_t)(c + 1)));
    if (nonnull) { o.push_back(8); gen_put_varint(o, gen_zigzag(v)); } else o.push_back(0);
  }
}
inline void gen_put_key(const b2_gen_spec& s, uint64_t handle, uint64_t commit_ts, uint8_t* k) {
  uint8_t raw[19];
  raw[0] = 't';
  uint64_t t = (uint64_t)s.table_id ^ 0x8000000000000000ull, h = handle ^ 0x8000000000000000ull;
  for (int i = 0; i < 8; ++i) { raw[1 + i] = (uint8_t)(t >> (8 * (7 - i))); raw[11 + i] = (uint8_t)(h >> (8 * (7 - i))); }
  raw[9] = '_'; raw[10] = 'r';
  for (int i = 0; i < 8; ++i) { k[i] = raw[i]; k[9 + i] = raw[8 + i]; }
  k[8] = 0xff; k[17] = 0xff;
  k[18] = raw[16]; k[19] = raw[17]; k[20] = raw[18];
  for (int i = 21; i < 26; ++i) k[i] = 0;
  k[26] = 0xff - 5;
  uint64_t nts = ~commit_ts;
  for (int i = 0; i < 8; ++i) k[27 + i] = (uint8_t)(nts >> (8 * (7 - i)));
}

// rows [first_handle, first_handle + n_rows) of the spec'ed table as one CF_WRITE block (heaps padded by 16 bytes)
inline void gen_block(const b2_gen_spec& s, GenBlock* out) {
  out->keys.clear(); out->vals.clear(); out->koff.clear(); out->voff.clear();
  out->keys.reserve((size_t)s.n_rows * 35 + 32);
  out->n_user_keys = s.n_rows;
  std::vector<uint8_t> row;
  for (uint64_t r = 0; r < s.n_rows; ++r) {
    const uint64_t handle = s.first_handle + r;
    const GenRowKind g = gen_row_kind(s, handle);
    uint64_t cts[3], sts[3], salt[3]; uint8_t typ[3]; bool with_row[3];
    int n = 0;  // versions newest first
    if (g.kind == 1) { cts[n] = s.newer_ts; typ[n] = 'P'; sts[n] = s.newer_ts - 1; salt[n] = 77; with_row[n] = true; ++n; }
    if (g.kind == 3) { cts[n] = s.commit_ts + 1; typ[n] = 'L'; sts[n] = s.commit_ts; salt[n] = 0; with_row[n] = false; ++n; }
    if (g.kind == 2) { cts[n] = s.commit_ts; typ[n] = 'D'; sts[n] = s.commit_ts - 1; salt[n] = 0; with_row[n] = false; ++n; }
    else { cts[n] = s.commit_ts; typ[n] = 'P'; sts[n] = s.commit_ts - 1; salt[n] = 0; with_row[n] = true; ++n; }
    if (g.kind == 1 || g.kind == 2) { cts[n] = s.commit_ts - 10; typ[n] = 'P'; sts[n] = s.commit_ts - 11; salt[n] = 99; with_row[n] = true; ++n; }
    for (int i = 0; i < n; ++i) {
      out->koff.push_back((uint32_t)out->keys.size());
      out->keys.resize(out->keys.size() + 35);
      gen_put_key(s, handle, cts[i], out->keys.data() + out->keys.size() - 35);
      out->voff.push_back((uint32_t)out->vals.size());
      out->vals.push_back(typ[i]);
      gen_put_varint(out->vals, sts[i]);
      if (with_row[i]) {
        row.clear();
        gen_row_bytes(s, handle, salt[i], row);
        out->vals.push_back('v'); out->vals.push_back((uint8_t)row.size());
        out->vals.insert(out->vals.end(), row.begin(), row.end());
      }
      if (typ[i] == 'L') {  // last_change -> the Put right below, 1 version away
        out->vals.push_back('l');
        for (int b = 0; b < 8; ++b) out->vals.push_back((uint8_t)(s.commit_ts >> (8 * (7 - b))));
        out->vals.push_back(1);
      }
    }
  }
  out->koff.push_back((uint32_t)out->keys.size());
  out->voff.push_back((uint32_t)out->vals.size());
  out->keys.resize(out->keys.size() + 32, 0);  // readable past the last entry (b2_copr.h: heaps are padded)
  out->vals.resize(out->vals.size() + 32, 0);
}

}  // namespace orc
