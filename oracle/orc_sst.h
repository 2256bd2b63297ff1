// TEST INFRASTRUCTURE — CPU oracle, not product code.
//
// RocksDB BlockBasedTable data blocks: builder and iterator, restated.  RocksDB is NOT under /root/reference: the
// reference reaches it through `rocksdb 0.3.0` / `librocksdb_sys 0.1.0` (git tikv/rust-rocksdb @ a6b0bb04,
// Cargo.lock:3844-3846, 5850-5852) behind components/engine_rocks; TiKV sets the block size and the table format
// version only (src/config/mod.rs:966 write CF 32 KiB, :704 format_version default 2; block_restart_interval keeps
// RocksDB's default 16).  What is restated is the published data-block layout (RocksDB `block_builder.cc` header
// comment / BlockBasedTable format wiki):
//
//   entry   := varint32 shared_bytes | varint32 unshared_bytes | varint32 value_length | key_delta | value
//   block   := entry* | fixed32 restarts[num_restarts] | fixed32 num_restarts     (restart entries have shared_bytes 0)
//   trailer := 1-byte compression type | fixed32 checksum                        (5 bytes, when the stored form is used)
//   key     := user key | fixed64 (sequence << 8 | value type)                    (internal key; kTypeValue = 1)
//   TiKV user key := 'z' | CF key (components/keys/src/lib.rs:28 DATA_PREFIX)
//
// PARITY UNPINNED for this file: the reference's tree holds no golden data-block bytes (blocks are produced and
// consumed inside librocksdb).  The builder and the iterator pin each other (round trip), the hand-assembled block in
// tests/test_sst_cpu.py pins both against the layout above, and the device encoder / decoder are compared with them.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace orc {

struct SstOptions {
  uint32_t restart_interval = 16;   // BlockBasedTableOptions::block_restart_interval
  uint32_t block_size = 32 * 1024;  // flush when the estimated block size reaches it (FlushBlockBySizePolicy); 0 = by entries_per_block
  uint32_t entries_per_block = 0;   // fixed entry count per block (the device encoder's policy) when block_size == 0
  uint32_t key_prefix_len = 1;      // bytes of `key_prefix_byte` put in front of every key
  uint8_t key_prefix_byte = 'z';
  uint32_t key_suffix_len = 8;      // 8 = append the internal-key footer (seq 0, kTypeValue)
  uint32_t trailer_len = 5;         // 5 = append the block trailer (type 0 = kNoCompression, checksum field zero), 0 = contents only
};

inline void sst_put_var32(std::string* dst, uint32_t v) {
  while (v >= 0x80) { dst->push_back((char)(v | 0x80)); v >>= 7; }
  dst->push_back((char)v);
}
inline void sst_put_fixed32(std::string* dst, uint32_t v) { for (int i = 0; i < 4; ++i) dst->push_back((char)(v >> (8 * i))); }

// BlockBuilder::Add / Finish
struct SstBlockBuilder {
  const SstOptions& o;
  std::string buf, last_key;
  std::vector<uint32_t> restarts{0};
  uint32_t counter = 0, n = 0;
  explicit SstBlockBuilder(const SstOptions& opt) : o(opt) {}
  size_t estimate() const { return buf.size() + restarts.size() * 4 + 4; }
  void add(const std::string& key, const uint8_t* val, uint32_t vlen) {
    uint32_t shared = 0;
    if (counter >= o.restart_interval) { restarts.push_back((uint32_t)buf.size()); counter = 0; }
    else if (n) {
      const size_t m = std::min(last_key.size(), key.size());
      while (shared < m && last_key[shared] == key[shared]) ++shared;
    }
    sst_put_var32(&buf, shared);
    sst_put_var32(&buf, (uint32_t)key.size() - shared);
    sst_put_var32(&buf, vlen);
    buf.append(key, shared, std::string::npos);
    buf.append((const char*)val, vlen);
    last_key = key;
    ++counter; ++n;
  }
  std::string finish() {
    for (uint32_t r : restarts) sst_put_fixed32(&buf, r);
    sst_put_fixed32(&buf, (uint32_t)restarts.size());
    for (uint32_t i = 0; i < o.trailer_len; ++i) buf.push_back(0);
    return buf;
  }
};

// flat sorted entries -> data blocks, back to back; block_offs gets n_blocks + 1 offsets
inline void sst_build(const uint8_t* keys, const uint32_t* koff, const uint8_t* vals, const uint32_t* voff, uint32_t n, const SstOptions& o, std::string* out,
                      std::vector<uint64_t>* block_offs) {
  out->clear(); block_offs->assign(1, 0);
  uint32_t e = 0;
  while (e < n) {
    SstBlockBuilder b(o);
    while (e < n) {
      std::string k(o.key_prefix_len, (char)o.key_prefix_byte);
      k.append((const char*)keys + koff[e], koff[e + 1] - koff[e]);
      if (o.key_suffix_len == 8) { k.push_back(1); k.append(7, '\0'); }  // fixed64 LE of (0 << 8 | kTypeValue)
      b.add(k, vals + voff[e], voff[e + 1] - voff[e]);
      ++e;
      if (o.block_size ? b.estimate() >= o.block_size : b.n >= o.entries_per_block) break;
    }
    out->append(b.finish());
    block_offs->push_back(out->size());
  }
}

inline bool sst_get_var32(const uint8_t*& p, const uint8_t* lim, uint32_t* v) {
  uint32_t r = 0;
  for (int i = 0; i < 5 && p < lim; ++i) {
    uint32_t b = *p++;
    r |= (b & 0x7f) << (7 * i);
    if (!(b & 0x80)) { *v = r; return true; }
  }
  return false;
}

// Block::Iter::SeekToFirst / Next over every block: 0 ok, 1 corrupted, 2 unsupported (hash index / value type != kTypeValue)
inline int sst_decode(const uint8_t* data, const uint64_t* block_offs, uint32_t n_blocks, uint32_t trailer_len, uint32_t key_prefix_len, uint32_t key_suffix_len,
                      std::string* keys, std::vector<uint32_t>* koff, std::string* vals, std::vector<uint32_t>* voff) {
  keys->clear(); vals->clear(); koff->assign(1, 0); voff->assign(1, 0);
  bool unsupported = false;
  for (uint32_t b = 0; b < n_blocks; ++b) {
    if (block_offs[b + 1] < block_offs[b] + trailer_len + 4) return 1;
    const uint8_t* base = data + block_offs[b];
    const uint32_t len = (uint32_t)(block_offs[b + 1] - block_offs[b]) - trailer_len;
    uint32_t nr;
    memcpy(&nr, base + len - 4, 4);
    if (nr >> 31) return 2;
    if (nr == 0 || (uint64_t)nr * 4 + 4 > len) return 1;
    const uint8_t *p = base, *lim = base + len - 4 - 4 * nr;
    // The restart array is what Seek lands on (Block::Iter::Seek -> SeekToRestartPoint + DecodeEntry, which reports a
    // corrupted block when a restart entry has shared != 0), and what the device decoder cuts its work at: every restart
    // offset must be the start of an entry with shared == 0, the first one 0, in ascending order.
    std::vector<uint32_t> rs(nr);
    memcpy(rs.data(), lim, 4ull * nr);
    if (rs[0] != 0) return 1;
    for (uint32_t j = 1; j < nr; ++j)
      if (rs[j] < rs[j - 1] || rs[j] > (uint32_t)(lim - base)) return 1;
    uint32_t next_restart = 0;
    std::string key;
    while (p < lim) {
      uint32_t sh, ns, vl;
      const uint32_t at = (uint32_t)(p - base);
      while (next_restart < nr && rs[next_restart] < at) return 1;  // a restart offset inside an entry
      if (!sst_get_var32(p, lim, &sh) || !sst_get_var32(p, lim, &ns) || !sst_get_var32(p, lim, &vl)) return 1;
      if (next_restart < nr && rs[next_restart] == at) {
        if (sh != 0) return 1;
        while (next_restart < nr && rs[next_restart] == at) ++next_restart;
      }
      if (sh > key.size() || (uint64_t)(lim - p) < (uint64_t)ns + vl) return 1;
      key.resize(sh);
      key.append((const char*)p, ns);
      p += ns;
      if (key.size() < (size_t)key_prefix_len + key_suffix_len) return 1;
      if (key_suffix_len == 8 && (uint8_t)key[key.size() - 8] != 1) unsupported = true;  // reported after the structure of every block was checked
      keys->append(key, key_prefix_len, key.size() - key_prefix_len - key_suffix_len);
      vals->append((const char*)p, vl);
      p += vl;
      koff->push_back((uint32_t)keys->size());
      voff->push_back((uint32_t)vals->size());
    }
    for (; next_restart < nr; ++next_restart)
      if (rs[next_restart] != (uint32_t)(lim - base)) return 1;  // (a restart offset at the very end: an empty last interval)
  }
  return unsupported ? 2 : 0;
}

}  // namespace orc
