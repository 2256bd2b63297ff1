// TEST INFRASTRUCTURE — CPU oracle, not product code.
// MVCC forward scan restatement:
//   Cursor                     components/tikv_kv/src/cursor.rs (seek / next / valid / stats)
//   WriteRef::parse            components/txn_types/src/write.rs:296-361, gc fence :425-442
//   Lock::parse / conflicts    components/txn_types/src/lock.rs:343-476, 520-611
//   ForwardScanner<LatestKv>   src/storage/mvcc/reader/scanner/forward.rs:76-109, 172-515
//   near_load_data_by_write    src/storage/mvcc/reader/scanner/mod.rs:371-402
#pragma once
#include <string>
#include <vector>

#include "../include/b2_copr.h"
#include "orc_codec.h"

namespace orc {

const uint64_t SEEK_BOUND = 8;  // components/tikv_kv/src/lib.rs:76

struct Error {
  int status = B2_OK;
  int mysql_code = 0;
  std::string msg;
  bool ok() const { return status == B2_OK; }
  static Error make(int st, const std::string& m, int code = 0) { Error e; e.status = st; e.msg = m; e.mysql_code = code; return e; }
};

// A column family = list of sorted blocks, viewed as one ordered sequence of entries.
struct CfView {
  std::vector<const b2_cf_block*> blocks;
  std::vector<uint64_t> base;  // global index of first entry of block i; base.back() = total
  void init(const b2_cf_block* b, uint32_t n) {
    blocks.clear(); base.assign(1, 0);
    for (uint32_t i = 0; i < n; ++i) { blocks.push_back(&b[i]); base.push_back(base.back() + b[i].n); }
  }
  uint64_t size() const { return base.back(); }
  void locate(uint64_t i, size_t* blk, uint32_t* off) const {
    size_t lo = 0, hi = blocks.size();
    while (hi - lo > 1) { size_t mid = (lo + hi) / 2; if (base[mid] <= i) lo = mid; else hi = mid; }
    *blk = lo; *off = (uint32_t)(i - base[lo]);
  }
  Slice key(uint64_t i) const {
    size_t b; uint32_t o; locate(i, &b, &o);
    const b2_cf_block* k = blocks[b];
    return Slice(k->keys + k->key_offs[o], k->key_offs[o + 1] - k->key_offs[o]);
  }
  Slice value(uint64_t i) const {
    size_t b; uint32_t o; locate(i, &b, &o);
    const b2_cf_block* k = blocks[b];
    return Slice(k->vals + k->val_offs[o], k->val_offs[o + 1] - k->val_offs[o]);
  }
  uint64_t lower_bound(Slice target) const {  // first entry with key >= target
    uint64_t lo = 0, hi = size();
    while (lo < hi) { uint64_t mid = lo + (hi - lo) / 2; if (cmp_bytes(key(mid), target) < 0) lo = mid + 1; else hi = mid; }
    return lo;
  }
};

struct CfStatistics {  // components/tikv_kv/src/stats.rs:20-40
  uint64_t processed_keys = 0, get = 0, next = 0, prev = 0, seek = 0, seek_for_prev = 0, over_seek_bound = 0;
};
struct Statistics {
  CfStatistics lock, write, data;
  uint64_t processed_size = 0;
};

// Cursor bounded to [lo, hi) of a CfView (the iterator's lower/upper bound).
struct Cursor {
  const CfView* cf = nullptr;
  uint64_t lo = 0, hi = 0, pos = 0;
  void init(const CfView* c, Slice lower, bool has_lower, Slice upper, bool has_upper) {
    cf = c;
    lo = has_lower ? c->lower_bound(lower) : 0;
    hi = has_upper ? c->lower_bound(upper) : c->size();
    if (hi < lo) hi = lo;
    pos = hi;
  }
  bool valid() const { return pos >= lo && pos < hi; }
  void seek(Slice key, CfStatistics* st) { st->seek++; uint64_t p = cf->lower_bound(key); pos = p < lo ? lo : p; }
  void seek_to_first(CfStatistics* st) { st->seek++; pos = lo; }
  void next(CfStatistics* st) { st->next++; pos++; }
  // backward movement (storage/kv/cursor.rs reverse_seek :277-336, seek_for_prev :338-383, seek_to_last :496-503, prev :533-545)
  void prev(CfStatistics* st) { st->prev++; pos--; }  // lo - 1 (or wrap-around below 0) is "not valid"
  void seek_to_last(CfStatistics* st) { st->seek++; pos = hi - 1; }
  void seek_for_prev(Slice key, CfStatistics* st) {  // last entry <= key
    st->seek_for_prev++;
    uint64_t p = cf->lower_bound(key);
    if (p < cf->size() && cmp_bytes(cf->key(p), key) == 0) ++p;
    if (p > hi) p = hi;
    pos = p - 1;
  }
  void reverse_seek(Slice key, CfStatistics* st) {  // last entry < key
    st->seek_for_prev++;
    uint64_t p = cf->lower_bound(key);
    if (p > hi) p = hi;
    pos = p - 1;
  }
  Slice key() const { return cf->key(pos); }
  Slice value() const { return cf->value(pos); }
};

// ---- write record ----
enum WriteType { WT_PUT, WT_DELETE, WT_LOCK, WT_ROLLBACK };
enum LastChangeKind { LC_UNKNOWN, LC_EXIST, LC_NOT_EXIST };
struct WriteRef {
  WriteType write_type;
  uint64_t start_ts = 0;
  bool has_short_value = false;
  Slice short_value;
  bool has_overlapped_rollback = false;
  bool has_gc_fence = false;
  uint64_t gc_fence = 0;
  LastChangeKind last_change = LC_UNKNOWN;
  uint64_t last_change_ts = 0, estimated_versions_to_last_change = 0;
  uint64_t txn_source = 0;
};

inline bool write_parse(Slice b, WriteRef* w, std::string* err) {
  if (b.empty()) { *err = "bad format write"; return false; }
  switch (b[0]) {
    case 'P': w->write_type = WT_PUT; break;
    case 'D': w->write_type = WT_DELETE; break;
    case 'L': w->write_type = WT_LOCK; break;
    case 'R': w->write_type = WT_ROLLBACK; break;
    default: *err = "bad format write"; return false;
  }
  b = b.sub(1);
  size_t n = decode_var_u64(b, &w->start_ts);
  if (!n) { *err = "bad format write"; return false; }
  b = b.sub(n);
  uint64_t lc_ts = 0, lc_ver = 0;
  while (!b.empty()) {
    uint8_t tag = b[0];
    b = b.sub(1);
    if (tag == 'v') {
      if (b.empty()) { *err = "bad format write"; return false; }
      size_t len = b[0];
      b = b.sub(1);
      if (b.n < len) { *err = "content len shorter than short value len (panic)"; return false; }
      w->has_short_value = true; w->short_value = b.sub(0, len);
      b = b.sub(len);
    } else if (tag == 'R') {
      w->has_overlapped_rollback = true;
    } else if (tag == 'F') {
      if (b.n < 8) { *err = "unexpected eof"; return false; }
      w->has_gc_fence = true; w->gc_fence = get_u64_be(b.p);
      b = b.sub(8);
    } else if (tag == 'l') {
      if (b.n < 8) { *err = "unexpected eof"; return false; }
      lc_ts = get_u64_be(b.p);
      b = b.sub(8);
      size_t m = decode_var_u64_tu(b, &lc_ver);
      if (!m) { *err = "unexpected eof"; return false; }
      b = b.sub(m);
    } else if (tag == 'S') {
      size_t m = decode_var_u64_tu(b, &w->txn_source);
      if (!m) { *err = "unexpected eof"; return false; }
      b = b.sub(m);
    } else {
      break;  // forward compatibility: unknown tag stops parsing (:344-348)
    }
  }
  // LastChange::from_parts (types.rs:721-731)
  if (lc_ts == 0) w->last_change = lc_ver > 0 ? LC_NOT_EXIST : LC_UNKNOWN;
  else {
    if (lc_ver == 0) { *err = "last_change versions must be > 0 (assert)"; return false; }
    w->last_change = LC_EXIST;
  }
  w->last_change_ts = lc_ts; w->estimated_versions_to_last_change = lc_ver;
  return true;
}
inline bool write_check_gc_fence_as_latest_version(const WriteRef& w, uint64_t read_ts) {  // :425-442
  if (w.has_gc_fence && w.gc_fence != 0 && w.gc_fence <= read_ts) return false;
  return true;
}

// ---- lock record (fields needed by check_ts_conflict_si) ----
struct LockRec {
  uint8_t lock_type = 0;  // 'P','D','L','S','H'
  Bytes primary;
  uint64_t ts = 0, ttl = 0, min_commit_ts = 0, for_update_ts = 0, txn_size = 0;
  bool use_async_commit = false, use_one_pc = false;
};
// compact bytes: var_i64 len + bytes (codec byte.rs:540-)
inline bool read_compact_bytes(Slice* b, Bytes* out) {
  int64_t len;
  size_t n = decode_var_i64(*b, &len);
  if (!n || len < 0 || (uint64_t)len > b->n - n) return false;
  out->assign(b->p + n, b->p + n + len);
  *b = b->sub(n + (size_t)len);
  return true;
}
inline bool lock_parse(Slice b, LockRec* l, std::string* err) {  // lock.rs:520-611
  if (b.empty()) { *err = "bad format lock"; return false; }
  uint8_t t = b[0];
  if (t != 'P' && t != 'D' && t != 'L' && t != 'S' && t != 'H') { *err = "bad format lock"; return false; }
  l->lock_type = t;
  if (t == 'H') return true;  // SharedLocks: ignored by check_ts_conflict_si (:353-356)
  b = b.sub(1);
  if (!read_compact_bytes(&b, &l->primary)) { *err = "bad format lock"; return false; }
  size_t n = decode_var_u64_tu(b, &l->ts);
  if (!n) { *err = "bad format lock"; return false; }
  b = b.sub(n);
  if (b.empty()) return true;
  n = decode_var_u64_tu(b, &l->ttl);
  if (!n) { *err = "bad format lock"; return false; }
  b = b.sub(n);
  while (!b.empty()) {
    uint8_t tag = b[0];
    b = b.sub(1);
    if (tag == 'v') {
      if (b.empty()) { *err = "bad format lock"; return false; }
      size_t len = b[0];
      if (b.n - 1 < len) { *err = "bad format lock"; return false; }
      b = b.sub(1 + len);
    } else if (tag == 'f' || tag == 't' || tag == 'c' || tag == 'g') {
      if (b.n < 8) { *err = "bad format lock"; return false; }
      uint64_t v = get_u64_be(b.p);
      if (tag == 'f') l->for_update_ts = v; else if (tag == 't') l->txn_size = v; else if (tag == 'c') l->min_commit_ts = v;
      b = b.sub(8);
    } else if (tag == 'a') {
      l->use_async_commit = true;
      uint64_t cnt;
      n = decode_var_u64_tu(b, &cnt);
      if (!n) { *err = "bad format lock"; return false; }
      b = b.sub(n);
      for (uint64_t i = 0; i < cnt; ++i) { Bytes tmp; if (!read_compact_bytes(&b, &tmp)) { *err = "bad format lock"; return false; } }
    } else if (tag == 'r') {
      uint64_t cnt;
      n = decode_var_u64_tu(b, &cnt);
      if (!n || b.n - n < cnt * 8) { *err = "bad format lock"; return false; }
      b = b.sub(n + cnt * 8);
    } else if (tag == 'l') {
      if (b.n < 8) { *err = "bad format lock"; return false; }
      b = b.sub(8);
      uint64_t v; n = decode_var_u64_tu(b, &v);
      if (!n) { *err = "bad format lock"; return false; }
      b = b.sub(n);
    } else if (tag == 's') {
      uint64_t v; n = decode_var_u64_tu(b, &v);
      if (!n) { *err = "bad format lock"; return false; }
      b = b.sub(n);
    } else if (tag == 'F') {
      // is_locked_with_conflict
    } else break;
  }
  return true;
}

struct ScannerConfig {  // scanner/mod.rs:264-289
  uint64_t ts = 0;
  int isolation_level = B2_ISO_SI;
  bool check_has_newer_ts_data = false;
  bool load_commit_ts = false;
  bool omit_value = false;
  std::vector<uint64_t> bypass_locks, access_locks;
  bool has_lower = false, has_upper = false;
  Bytes lower_bound, upper_bound;  // encoded user keys
};

inline bool ts_set_contains(const std::vector<uint64_t>& s, uint64_t ts) {
  for (uint64_t x : s) if (x == ts) return true;
  return false;
}

// lock.rs:343-416 check_ts_conflict_si (is_replica_read = false). returns true if conflict.
inline bool check_ts_conflict_si(const LockRec& lock, Slice user_key_encoded, uint64_t ts, const std::vector<uint64_t>& bypass) {
  if (lock.lock_type == 'H') return false;
  if (lock.ts > ts || lock.lock_type == 'L' || lock.lock_type == 'S') return false;
  if (lock.min_commit_ts > ts) return false;
  if (ts_set_contains(bypass, lock.ts)) return false;
  if (ts == ~0ull) {
    Bytes raw;
    decode_bytes(user_key_encoded, &raw);
    if (raw == lock.primary && !lock.use_async_commit && !lock.use_one_pc) return false;
  }
  return true;
}

struct ScanOutput {
  Bytes user_key;  // encoded user key (without ts)
  Bytes value;
  bool has_commit_ts = false;
  uint64_t commit_ts = 0;
};

enum NewerTsCheckState { NEWER_UNKNOWN = -1, NEWER_NOT_MET = 0, NEWER_MET = 1 };

struct ForwardScanner {
  ScannerConfig cfg;
  const CfView* write_cf = nullptr; const CfView* lock_cf = nullptr; const CfView* default_cf = nullptr;
  Cursor write, lock, dflt;
  bool has_lock_cursor = false, dflt_ready = false;
  bool is_started = false;
  Bytes cur_key_buf_;
  Statistics statistics;
  int met_newer_ts_data = NEWER_UNKNOWN;

  void init(const ScannerConfig& c, const CfView* w, const CfView* l, const CfView* d) {
    cfg = c; write_cf = w; lock_cf = l; default_cf = d;
    Slice lo(cfg.lower_bound.data(), cfg.lower_bound.size()), hi(cfg.upper_bound.data(), cfg.upper_bound.size());
    write.init(w, lo, cfg.has_lower, hi, cfg.has_upper);
    // scanner/mod.rs:206-214: lock cursor only created for SI / RcCheckTs
    has_lock_cursor = l && l->size() > 0 && cfg.isolation_level != B2_ISO_RC;
    if (has_lock_cursor) lock.init(l, lo, cfg.has_lower, hi, cfg.has_upper);
    met_newer_ts_data = cfg.check_has_newer_ts_data ? NEWER_NOT_MET : NEWER_UNKNOWN;
    is_started = false; dflt_ready = false;
  }

  // forward.rs:76-109
  void move_write_cursor_to_next_user_key(const Bytes& current_user_key) {
    Slice uk(current_user_key.data(), current_user_key.size());
    for (uint64_t i = 0; i < SEEK_BOUND; ++i) {
      if (i > 0) write.next(&statistics.write);
      if (!write.valid()) return;
      if (!is_user_key_eq(write.key(), uk)) return;
    }
    statistics.write.over_seek_bound++;
    Bytes k = key_append_ts(current_user_key, 0);
    write.seek(Slice(k.data(), k.size()), &statistics.write);
  }

  // forward.rs:310-375. returns 1 = still on user key at a version <= ts, 0 = moved off, -1 = error
  int move_write_cursor_to_ts(const Bytes& user_key, Error* err) {
    Slice uk(user_key.data(), user_key.size());
    for (uint64_t i = 0; i < SEEK_BOUND; ++i) {
      if (i > 0) { write.next(&statistics.write); if (!write.valid()) return 0; }
      Slice ck = write.key();
      if (!is_user_key_eq(ck, uk)) return 0;
      uint64_t commit_ts = decode_u64_desc(ck.p + ck.n - 8);
      if (commit_ts <= cfg.ts) return 1;
      if (met_newer_ts_data == NEWER_NOT_MET) met_newer_ts_data = NEWER_MET;
      if (cfg.isolation_level == B2_ISO_RC_CHECK_TS) {
        *err = Error::make(B2_ERR_WRITE_CONFLICT, "write conflict (RcCheckTs): newer version exists");
        return -1;
      }
    }
    statistics.write.over_seek_bound++;
    Bytes k = key_append_ts(user_key, cfg.ts);
    write.seek(Slice(k.data(), k.size()), &statistics.write);
    if (!write.valid()) return 0;
    if (!is_user_key_eq(write.key(), uk)) return 0;
    return 1;
  }

  // scanner/mod.rs:371-402 near_load_data_by_write
  bool load_default(const Bytes& user_key, uint64_t start_ts, Bytes* out, Error* err) {
    if (!dflt_ready) { dflt.init(default_cf, Slice(), false, Slice(), false); dflt_ready = true; }
    Bytes seek_key = key_append_ts(user_key, start_ts);
    Slice sk(seek_key.data(), seek_key.size());
    statistics.data.seek++;
    uint64_t p = default_cf ? default_cf->lower_bound(sk) : 0;
    if (!default_cf || p >= default_cf->size() || cmp_bytes(default_cf->key(p), sk) != 0) {
      *err = Error::make(B2_ERR_STORAGE, "default not found");
      return false;
    }
    statistics.data.processed_keys++;
    Slice v = default_cf->value(p);
    out->assign(v.p, v.p + v.n);
    return true;
  }

  // LatestKvPolicy::handle_write forward.rs:433-515.  returns 1 = output filled, 0 = skip, -1 = error
  int handle_write(const Bytes& current_user_key, ScanOutput* out, Error* err) {
    Slice uk(current_user_key.data(), current_user_key.size());
    bool has_value = false;
    for (;;) {
      WriteRef w;
      std::string perr;
      if (!write_parse(write.value(), &w, &perr)) { *err = Error::make(B2_ERR_STORAGE, perr); return -1; }
      if (!write_check_gc_fence_as_latest_version(w, cfg.ts)) break;
      if (w.write_type == WT_PUT) {
        out->has_commit_ts = cfg.load_commit_ts;
        if (cfg.load_commit_ts) { Slice k = write.key(); out->commit_ts = decode_u64_desc(k.p + k.n - 8); }
        if (cfg.omit_value) { out->value.clear(); has_value = true; break; }
        if (w.has_short_value) { out->value.assign(w.short_value.p, w.short_value.p + w.short_value.n); has_value = true; break; }
        if (!load_default(current_user_key, w.start_ts, &out->value, err)) return -1;
        has_value = true;
        break;
      } else if (w.write_type == WT_DELETE) {
        break;
      } else {
        if (w.last_change == LC_NOT_EXIST) break;
        if (w.last_change == LC_EXIST && w.estimated_versions_to_last_change >= SEEK_BOUND) {
          Bytes k = key_append_ts(current_user_key, w.last_change_ts);
          write.seek(Slice(k.data(), k.size()), &statistics.write);
        } else {
          write.next(&statistics.write);
        }
      }
      if (!write.valid()) return 0;
      if (!is_user_key_eq(write.key(), uk)) return 0;
    }
    move_write_cursor_to_next_user_key(current_user_key);
    if (has_value) { out->user_key = current_user_key; return 1; }
    return 0;
  }

  // read_next forward.rs:172-304. returns 1 = row, 0 = drained, -1 = error
  int read_next(ScanOutput* out, Error* err) {
    if (!is_started) {
      if (cfg.has_lower) {
        Slice lb(cfg.lower_bound.data(), cfg.lower_bound.size());
        write.seek(lb, &statistics.write);
        if (has_lock_cursor) lock.seek(lb, &statistics.lock);
      } else {
        write.seek_to_first(&statistics.write);
        if (has_lock_cursor) lock.seek_to_first(&statistics.lock);
      }
      is_started = true;
    }
    for (;;) {
      bool wv = write.valid();
      bool lv = has_lock_cursor && lock.valid();
      Bytes& current_user_key = cur_key_buf_;  // buffer reused across rows (Key::from_encoded_slice reserves once)
      current_user_key.clear();
      bool has_write, has_lock;
      if (!wv && !lv) return 0;
      if (!wv) { Slice lk = lock.key(); current_user_key.assign(lk.p, lk.p + lk.n); has_write = false; has_lock = true; }
      else {
        Slice wk = write.key();
        if (wk.n < 8) { *err = Error::make(B2_ERR_STORAGE, "key too short to truncate ts"); return -1; }
        Slice wuk(wk.p, wk.n - 8);
        if (!lv) { current_user_key.assign(wuk.p, wuk.p + wuk.n); has_write = true; has_lock = false; }
        else {
          Slice lk = lock.key();
          int c = cmp_bytes(wuk, lk);
          if (c < 0) { current_user_key.assign(wuk.p, wuk.p + wuk.n); has_write = true; has_lock = false; }
          else if (c > 0) { current_user_key.assign(lk.p, lk.p + lk.n); has_write = false; has_lock = true; }
          else { current_user_key.assign(lk.p, lk.p + lk.n); has_write = true; has_lock = true; }
        }
      }
      if (has_lock) {
        if (met_newer_ts_data == NEWER_NOT_MET) met_newer_ts_data = NEWER_MET;
        // LatestKvPolicy::handle_lock forward.rs:384-431 (SI; RcCheckTs treated via same path)
        LockRec lrec; std::string perr;
        if (!lock_parse(lock.value(), &lrec, &perr)) { *err = Error::make(B2_ERR_STORAGE, perr); return -1; }
        lock.next(&statistics.lock);
        bool conflict = false;
        Slice uk(current_user_key.data(), current_user_key.size());
        if (cfg.isolation_level == B2_ISO_SI) conflict = check_ts_conflict_si(lrec, uk, cfg.ts, cfg.bypass_locks);
        else if (cfg.isolation_level == B2_ISO_RC_CHECK_TS) {
          // lock.rs:418-455 check_ts_conflict_rc_check_ts
          conflict = !(lrec.lock_type == 'H' || lrec.lock_type == 'L' || lrec.lock_type == 'S' || ts_set_contains(cfg.bypass_locks, lrec.ts));
          if (conflict) { statistics.lock.processed_keys++; *err = Error::make(B2_ERR_WRITE_CONFLICT, "write conflict (RcCheckTs): lock"); return -1; }
        }
        if (conflict) {
          statistics.lock.processed_keys++;
          move_write_cursor_to_next_user_key(current_user_key);
          if (!cfg.load_commit_ts && ts_set_contains(cfg.access_locks, lrec.ts)) {
            *err = Error::make(B2_ERR_UNSUPPORTED, "access_locks read-through is not restated");
            return -1;
          }
          *err = Error::make(B2_ERR_KEY_IS_LOCKED, "key is locked, lock_version=" + std::to_string(lrec.ts));
          return -1;
        }
      }
      if (has_write) {
        int r = move_write_cursor_to_ts(current_user_key, err);
        if (r < 0) return -1;
        if (r == 1) {
          int h = handle_write(current_user_key, out, err);
          if (h < 0) return -1;
          if (h == 1) {
            statistics.write.processed_keys++;
            statistics.processed_size += out->user_key.size() + out->value.size();
            return 1;
          }
        }
      }
    }
  }
};

// ---- backward scan: src/storage/mvcc/reader/scanner/backward.rs ----------------------------------------------
// Same visible version per user key as the forward scanner (the newest Put / Delete with commit_ts <= ts, Lock and
// Rollback records skipped, gc fence checked on it), keys delivered in descending order.  Test infrastructure for the
// `desc` table scan; cursor statistics follow the reference's prev / seek pattern.
static const uint64_t REVERSE_SEEK_BOUND = 16;  // backward.rs:23

struct BackwardScanner {
  ScannerConfig cfg;
  const CfView* write_cf = nullptr; const CfView* lock_cf = nullptr; const CfView* default_cf = nullptr;
  Cursor write, lock;
  bool has_lock_cursor = false;
  bool is_started = false;
  Statistics statistics;
  int met_newer_ts_data = NEWER_UNKNOWN;

  void init(const ScannerConfig& c, const CfView* w, const CfView* l, const CfView* d) {
    cfg = c; write_cf = w; lock_cf = l; default_cf = d;
    Slice lo(cfg.lower_bound.data(), cfg.lower_bound.size()), hi(cfg.upper_bound.data(), cfg.upper_bound.size());
    write.init(w, lo, cfg.has_lower, hi, cfg.has_upper);
    has_lock_cursor = l && l->size() > 0 && cfg.isolation_level != B2_ISO_RC;
    if (has_lock_cursor) lock.init(l, lo, cfg.has_lower, hi, cfg.has_upper);
    met_newer_ts_data = cfg.check_has_newer_ts_data ? NEWER_NOT_MET : NEWER_UNKNOWN;
    is_started = false;
  }

  // scanner/mod.rs near_reverse_load_data_by_write: the value of (user_key, start_ts) in CF_DEFAULT
  bool load_default(const Bytes& user_key, uint64_t start_ts, Bytes* out, Error* err) {
    Bytes seek_key = key_append_ts(user_key, start_ts);
    uint64_t p = default_cf ? default_cf->lower_bound(Slice(seek_key.data(), seek_key.size())) : 0;
    statistics.data.seek_for_prev++;
    if (!default_cf || p >= default_cf->size() || cmp_bytes(default_cf->key(p), Slice(seek_key.data(), seek_key.size())) != 0) {
      *err = Error::make(B2_ERR_STORAGE, "default not found");
      return false;
    }
    statistics.data.processed_keys++;
    Slice v = default_cf->value(p);
    out->assign(v.p, v.p + v.n);
    return true;
  }

  // backward.rs:460-491
  void move_write_cursor_to_prev_user_key(const Bytes& current_user_key) {
    Slice uk(current_user_key.data(), current_user_key.size());
    for (uint64_t i = 0; i < SEEK_BOUND; ++i) {
      if (i > 0) write.prev(&statistics.write);
      if (!write.valid()) return;
      if (!is_user_key_eq(write.key(), uk)) return;
    }
    statistics.write.over_seek_bound++;
    write.seek_for_prev(uk, &statistics.write);  // the bare user key sorts before every version of it
  }

  // handle_last_version backward.rs:411-433 + reverse_load_data_by_write :440-458.  1 = value, 0 = none, -1 = error
  int finish(bool have, const WriteRef& w, const Bytes& short_value, uint64_t commit_ts, const Bytes& user_key, ScanOutput* out, Error* err) {
    if (!have) return 0;
    if (!write_check_gc_fence_as_latest_version(w, cfg.ts)) return 0;
    if (w.write_type == WT_DELETE) return 0;
    out->has_commit_ts = cfg.load_commit_ts; out->commit_ts = commit_ts;
    if (cfg.omit_value) { out->value.clear(); return 1; }
    if (w.has_short_value) { out->value = short_value; return 1; }
    return load_default(user_key, w.start_ts, &out->value, err) ? 1 : -1;
  }

  // reverse_get backward.rs:218-405: the write cursor points at the earliest version of user_key
  int reverse_get(const Bytes& user_key, bool* met_prev_user_key, ScanOutput* out, Error* err) {
    Slice uk(user_key.data(), user_key.size());
    bool have = false; WriteRef last; Bytes last_short; uint64_t loaded_commit_ts = 0, last_checked_commit_ts = 0;
    for (uint64_t i = 0; i < REVERSE_SEEK_BOUND; ++i) {
      if (i > 0) {
        write.prev(&statistics.write);
        if (!write.valid()) return finish(have, last, last_short, loaded_commit_ts, user_key, out, err);
      }
      Slice ck = write.key();
      last_checked_commit_ts = decode_u64_desc(ck.p + ck.n - 8);
      bool is_done = false;
      if (!is_user_key_eq(ck, uk)) { *met_prev_user_key = true; is_done = true; }
      else if (last_checked_commit_ts > cfg.ts) {
        is_done = true;
        if (met_newer_ts_data == NEWER_NOT_MET) met_newer_ts_data = NEWER_MET;
        if (cfg.isolation_level == B2_ISO_RC_CHECK_TS) { *err = Error::make(B2_ERR_WRITE_CONFLICT, "write conflict (RcCheckTs): newer version exists"); return -1; }
      }
      if (is_done) return finish(have, last, last_short, loaded_commit_ts, user_key, out, err);
      WriteRef w; std::string perr;
      if (!write_parse(write.value(), &w, &perr)) { *err = Error::make(B2_ERR_STORAGE, perr); return -1; }
      if (w.write_type == WT_PUT || w.write_type == WT_DELETE) {
        have = true; last = w;
        last_short.assign(w.short_value.p, w.short_value.p + (w.has_short_value ? w.short_value.n : 0));
        loaded_commit_ts = last_checked_commit_ts;
      }
    }
    if (last_checked_commit_ts == cfg.ts) {  // :292-307
      if (met_newer_ts_data == NEWER_NOT_MET) {
        write.prev(&statistics.write);
        if (write.valid()) {
          if (is_user_key_eq(write.key(), uk)) met_newer_ts_data = NEWER_MET; else *met_prev_user_key = true;
        }
      }
      return finish(have, last, last_short, loaded_commit_ts, user_key, out, err);
    }
    // many versions: seek to (user_key, ts) and walk forward to the newest Put / Delete below it (:314-404)
    if (met_newer_ts_data == NEWER_NOT_MET) {
      Bytes k = key_append_ts(user_key, ~0ull);
      write.seek(Slice(k.data(), k.size()), &statistics.write);
      Slice ck = write.key();
      if (decode_u64_desc(ck.p + ck.n - 8) > cfg.ts) met_newer_ts_data = NEWER_MET;
    }
    Bytes k = key_append_ts(user_key, cfg.ts);
    write.seek(Slice(k.data(), k.size()), &statistics.write);
    for (;;) {
      Slice ck = write.key();
      uint64_t current_ts = decode_u64_desc(ck.p + ck.n - 8);
      if (current_ts <= last_checked_commit_ts) return finish(have, last, last_short, loaded_commit_ts, user_key, out, err);
      WriteRef w; std::string perr;
      if (!write_parse(write.value(), &w, &perr)) { *err = Error::make(B2_ERR_STORAGE, perr); return -1; }
      if (!write_check_gc_fence_as_latest_version(w, cfg.ts)) return 0;
      if (w.write_type == WT_PUT) {
        Bytes sv(w.short_value.p, w.short_value.p + (w.has_short_value ? w.short_value.n : 0));
        out->has_commit_ts = cfg.load_commit_ts; out->commit_ts = current_ts;
        if (cfg.omit_value) { out->value.clear(); return 1; }
        if (w.has_short_value) { out->value = sv; return 1; }
        return load_default(user_key, w.start_ts, &out->value, err) ? 1 : -1;
      }
      if (w.write_type == WT_DELETE) return 0;
      write.next(&statistics.write);  // Lock / Rollback: next (older) version
    }
  }

  // read_next backward.rs:78-216.  1 = row, 0 = drained, -1 = error
  int read_next(ScanOutput* out, Error* err) {
    if (!is_started) {
      if (cfg.has_upper) {
        Slice ub(cfg.upper_bound.data(), cfg.upper_bound.size());
        write.reverse_seek(ub, &statistics.write);
        if (has_lock_cursor) lock.reverse_seek(ub, &statistics.lock);
      } else {
        write.seek_to_last(&statistics.write);
        if (has_lock_cursor) lock.seek_to_last(&statistics.lock);
      }
      is_started = true;
    }
    for (;;) {
      bool wv = write.valid(), lv = has_lock_cursor && lock.valid();
      if (!wv && !lv) return 0;
      Bytes current_user_key;
      bool has_write, has_lock;
      if (!wv) { Slice lk = lock.key(); current_user_key.assign(lk.p, lk.p + lk.n); has_write = false; has_lock = true; }
      else {
        Slice wk = write.key();
        if (wk.n < 8) { *err = Error::make(B2_ERR_STORAGE, "key too short to truncate ts"); return -1; }
        Slice wuk(wk.p, wk.n - 8);
        if (!lv) { current_user_key.assign(wuk.p, wuk.p + wuk.n); has_write = true; has_lock = false; }
        else {
          Slice lk = lock.key();
          int c = cmp_bytes(wuk, lk);  // descending: the larger key comes first
          if (c < 0) { current_user_key.assign(lk.p, lk.p + lk.n); has_write = false; has_lock = true; }
          else if (c > 0) { current_user_key.assign(wuk.p, wuk.p + wuk.n); has_write = true; has_lock = false; }
          else { current_user_key.assign(wuk.p, wuk.p + wuk.n); has_write = true; has_lock = true; }
        }
      }
      bool met_prev_user_key = false;
      if (has_lock) {
        LockRec lrec; std::string perr;
        if (!lock_parse(lock.value(), &lrec, &perr)) { *err = Error::make(B2_ERR_STORAGE, perr); return -1; }
        if (met_newer_ts_data == NEWER_NOT_MET) met_newer_ts_data = NEWER_MET;
        Slice uk(current_user_key.data(), current_user_key.size());
        bool conflict = false;
        if (cfg.isolation_level == B2_ISO_SI) conflict = check_ts_conflict_si(lrec, uk, cfg.ts, cfg.bypass_locks);
        else if (cfg.isolation_level == B2_ISO_RC_CHECK_TS)
          conflict = !(lrec.lock_type == 'H' || lrec.lock_type == 'L' || lrec.lock_type == 'S' || ts_set_contains(cfg.bypass_locks, lrec.ts));
        lock.prev(&statistics.lock);
        if (conflict) {
          statistics.lock.processed_keys++;
          if (cfg.isolation_level == B2_ISO_RC_CHECK_TS) { *err = Error::make(B2_ERR_WRITE_CONFLICT, "write conflict (RcCheckTs): lock"); return -1; }
          if (!cfg.load_commit_ts && ts_set_contains(cfg.access_locks, lrec.ts)) { *err = Error::make(B2_ERR_UNSUPPORTED, "access_locks read-through is not restated"); return -1; }
          if (has_write) move_write_cursor_to_prev_user_key(current_user_key);
          *err = Error::make(B2_ERR_KEY_IS_LOCKED, "key is locked, lock_version=" + std::to_string(lrec.ts));
          return -1;
        }
      }
      if (has_write) {
        int r = reverse_get(current_user_key, &met_prev_user_key, out, err);
        if (r < 0) return -1;
        if (!met_prev_user_key) move_write_cursor_to_prev_user_key(current_user_key);
        if (r == 1) {
          out->user_key = current_user_key;
          statistics.write.processed_keys++;
          statistics.processed_size += out->user_key.size() + out->value.size();
          return 1;
        }
      }
    }
  }
};

}  // namespace orc
