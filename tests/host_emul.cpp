// TEST HARNESS (not product, not oracle): drives the *device* row logic of tikv_b200/csrc/b2_device.h and the plan
// lowering of plan_compile.h on the CPU, entry by entry, the way scan_kernel does on the GPU.  There is no GPU in the
// authoring container; this lets `pytest -m "not gpu"` check the MVCC / row-decode / RPN / accumulator code that the
// kernels execute against the oracle before any GPU time is spent.  Kernel-only mechanics (tickets, look-back,
// shared-memory tables, atomics) are covered by the `-m gpu` tests.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../tikv_b200/csrc/plan_compile.h"

using namespace b2;

struct emu_result {
  int status = 0;
  int dev_err = 0;
  uint64_t err_entry = ~0ull;
  std::string msg;
  std::vector<int> kinds;
  std::vector<std::vector<uint64_t>> data;      // per output column: bits (i64 / f64)
  std::vector<std::vector<b2_decimal>> dec;     // per output column: decimals (when kind == DECIMAL)
  std::vector<std::vector<uint8_t>> nonnull;
  uint64_t n_rows = 0;
  uint64_t processed_keys = 0, processed_size = 0, met_newer = 0, dflt = 0;
  uint64_t checksum = 0, total_kvs = 0, total_bytes = 0;
};

static uint32_t lower_bound_block(const b2_cf_block& B, const std::vector<uint8_t>& key) {
  uint32_t lo = 0, hi = B.n;
  while (lo < hi) {
    uint32_t mid = lo + (hi - lo) / 2;
    if (bytes_cmp(B.keys + B.key_offs[mid], B.key_offs[mid + 1] - B.key_offs[mid], key.data(), (uint32_t)key.size()) < 0) lo = mid + 1; else hi = mid;
  }
  return lo;
}

static BlockView view_of(const b2_cf_block& c) { BlockView v; v.keys = c.keys; v.koff = c.key_offs; v.vals = c.vals; v.voff = c.val_offs; v.n = c.n; return v; }

struct GroupAcc { uint64_t w[MAX_ACC_WORDS]; };

// fast_kernel.cuh restated for the CPU: the branch-free front end of the lean kernels run over 32 consecutive entries at
// a time ("lanes" of a warp), with the run-ownership rules of b2_device.h fast_lane_decide.  For every entry it yields
// FA_COMMIT (the lane holds the visible Put of a plain run: row at val + row_off) and / or FA_PUSH (entry `e - push_back`
// goes to the general walk).  newer = the entry's commit_ts is above the snapshot.
struct LaneOut { uint32_t flags, push_back, row_off, row_len; bool newer; KeyTail tail; };
static void fast_front_warp(const BlockView& b, uint32_t e0, uint32_t e_lo, uint32_t e_hi, uint64_t read_ts, int isolation, LaneOut* out) {
  bool valid[32], same[32], vis[32], chosen[32], kok[32];
  uint32_t kind[32];
  KeyTail t[32];
  uint32_t start_m = 0, chosen_m = 0, valid_m = 0;
  for (uint32_t l = 0; l < 32; ++l) {
    const uint32_t e = e0 + l;
    valid[l] = e < e_hi;
    same[l] = vis[l] = chosen[l] = kok[l] = false; kind[l] = 0;
    out[l] = LaneOut{};
    if (!valid[l]) continue;
    valid_m |= 1u << l;
    const bool k35 = b.klen(e) == 35;
    kok[l] = k35 && key_tail_load(b.kptr(e), &t[l]);
    bool pvis = false;
    if (k35 && e != e_lo && b.klen(e - 1) == 35) {
      KeyTail q;
      key_tail_load(b.kptr(e - 1), &q);
      same[l] = key_tail_same_exact(t[l].a, t[l].b, q.a, q.b);
      pvis = key_tail_commit_ts(q) <= read_ts;
    }
    vis[l] = k35 && key_tail_commit_ts(t[l]) <= read_ts;
    chosen[l] = vis[l] && (!same[l] || !pvis);
    if (b.vlen(e) >= 2) kind[l] = fast_write_kind(b.vptr(e), b.vlen(e), &out[l].row_off, &out[l].row_len);
    if (!same[l]) start_m |= 1u << l;
    if (chosen[l]) chosen_m |= 1u << l;
    out[l].newer = k35 && !vis[l];
    out[l].tail = t[l];
  }
  for (uint32_t l = 0; l < 32; ++l)
    out[l].flags = fast_lane_decide(l, start_m, chosen_m, valid_m, valid[l], same[l], chosen[l], kok[l], kind[l], vis[l], isolation == B2_ISO_RC_CHECK_TS, &out[l].push_back);
}

static bool unit_prefix_ok(const BlockView& b, uint32_t lo, uint32_t hi) {  // kernels.cu unit_prefix_kernel
  if (hi <= lo || b.klen(lo) < 12 || b.klen(hi - 1) < 12) return false;
  return record_key_prefix_ok(b.kptr(lo)) && memcmp(b.kptr(lo), b.kptr(hi - 1), 12) == 0;
}
static bool g_emu_fast_front = true;
static uint64_t g_emu_fast_hits = 0;

extern "C" {

emu_result* emu_dag_handle(const b2_dag_plan* plan, const b2_key_range* ranges, uint32_t n_ranges, const b2_region_source* src) {
  emu_result* R = new emu_result();
  CompiledPlan cp;
  R->status = compile_plan(plan, &cp, &R->msg);
  if (R->status) return R;
  patch_pool_imms(cp, cp.pool.data());  // bytes constants: cell references into the host copy of the pool
  DevPlan& P = cp.dev;
  P.read_ts = src->read_ts; P.isolation = src->isolation_level;
  std::vector<BlockView> dviews;
  for (uint32_t i = 0; src->dflt && i < src->n_dflt; ++i) dviews.push_back(view_of(src->dflt[i]));
  DefaultCf dflt; dflt.blocks = dviews.data(); dflt.n_blocks = (uint32_t)dviews.size();

  size_t n_out = P.mode == PM_SCAN ? (size_t)P.n_out : cp.output_offsets.size();
  R->data.resize(n_out); R->dec.resize(n_out); R->nonnull.resize(n_out); R->kinds.resize(n_out);
  struct TopRow { TopItem it; std::vector<uint64_t> bits; std::vector<uint8_t> nn; };
  std::vector<TopRow> top_rows;
  typedef std::vector<uint64_t> GKey;  // one group-by expression: (is_null, key bits); several: value words then the NULL mask
  std::map<GKey, GroupAcc> groups;
  std::vector<GKey> group_order;
  GroupAcc single; memset(&single, 0, sizeof(single));
  uint64_t live_rows = 0;
  uint64_t err = ~0ull;
  auto report = [&](uint64_t entry, int code) { uint64_t v = (entry << 8) | (unsigned)code; if (v < err) err = v; };

  uint64_t base = 0;
  std::vector<uint64_t> bases;
  for (uint32_t b = 0; b < src->n_write; ++b) { bases.push_back(base); base += src->write[b].n; }
  for (uint32_t r = 0; r < n_ranges; ++r) {
    std::vector<uint8_t> lo = encode_memcomparable(ranges[r].start, ranges[r].start_len), hi = encode_memcomparable(ranges[r].end, ranges[r].end_len);
    for (uint32_t b = 0; b < src->n_write; ++b) {
      BlockView blk = view_of(src->write[b]);
      uint32_t e_lo = lower_bound_block(src->write[b], lo), e_hi = lower_bound_block(src->write[b], hi);
      const bool unit_fast = g_emu_fast_front && P.fast_n > 0 && unit_prefix_ok(blk, e_lo, e_hi);
      // pass 1 (lean kernel): per 32-entry warp, commit plain runs, push the others; pass 2 (general kernel, list mode):
      // the pushed entries.  Without the fast front end every entry is "pushed".
      std::vector<uint32_t> todo;                       // entries for the general walk, ascending
      std::vector<std::pair<uint32_t, LaneOut>> fast;   // committed lanes
      if (unit_fast) {
        for (uint32_t e0 = e_lo; e0 < e_hi; e0 += 32) {
          LaneOut lo[32];
          fast_front_warp(blk, e0, e_lo, e_hi, P.read_ts, P.isolation, lo);
          for (uint32_t l = 0; l < 32 && e0 + l < e_hi; ++l) {
            if (lo[l].newer) R->met_newer = 1;
            if (lo[l].flags & FA_PUSH) todo.push_back(e0 + l - ((lo[l].flags & FA_COMMIT) ? 0 : lo[l].push_back));
            if ((lo[l].flags & FA_COMMIT) && !(lo[l].flags & FA_PUSH)) fast.push_back({e0 + l, lo[l]});
          }
        }
      } else {
        for (uint32_t e = e_lo; e < e_hi; ++e) todo.push_back(e);
      }
      // merge both streams in entry order (the scan output is ordered)
      std::sort(todo.begin(), todo.end());
      size_t fi = 0, ti = 0;
      while (fi < fast.size() || ti < todo.size()) {
        Row row; Cells cells;
        uint8_t idx_buf[IDX_RAW_MAX];
        row.imms = cp.imms;
        bool keep = false;
        uint32_t e;
        bool take_fast = ti >= todo.size() || (fi < fast.size() && fast[fi].first < todo[ti]);
        if (take_fast) {
          e = fast[fi].first;
          const LaneOut& lo = fast[fi].second;
          ++fi;
          row.enc_key = blk.kptr(e); row.enc_key_len = 27; row.commit_ts = key_tail_commit_ts(lo.tail);
          bool ok = fast_row_v2(P, blk.vptr(e) + lo.row_off, lo.row_len, row);
          row.gv = blk.vptr(e) + lo.row_off;
          if (ok) { row.filled = P.fast_filled; ok = eval_conds(P, row, cells, &keep) == 0; }
          if (!ok) {  // the row needs the general decoder: the lane pushes its run start instead of committing
            todo.insert(std::upper_bound(todo.begin() + ti, todo.end(), e - lo.push_back), e - lo.push_back);
            continue;
          }
          g_emu_fast_hits++;
          R->processed_keys++; R->processed_size += 27 + lo.row_len;
        } else {
          e = todo[ti++];
          if (ti < todo.size() + 1 && ti >= 2 && todo[ti - 2] == e) { R->status = B2_ERR_INVALID_ARG; R->msg = "emulation: run start pushed twice"; }
        bool start = (e == e_lo) || !same_user_key(blk, e - 1, e);
        if (!start) continue;
        RunOut ro;
        resolve_run(blk, e, e_hi, e_hi, P.read_ts, P.isolation, dflt, &ro);
        R->met_newer |= ro.met_newer; R->dflt += ro.dflt_lookup;
        if (ro.err) { report(bases[b] + e, ro.err); continue; }
        if (!ro.found) continue;
        uint32_t ko = blk.koff[e], kl = blk.koff[e + 1] - ko;
        R->processed_keys++; R->processed_size += (kl - 8) + ro.val_len;
        row.enc_key = blk.keys + ko; row.enc_key_len = kl - 8; row.commit_ts = ro.commit_ts;
        int er;
        if (P.idx_cols > 0) er = index_row_split(P, row, cells, ro.val, ro.val_len, idx_buf);
        else { er = row_open(ro.val, ro.val_len, &row.rv); row.gv = ro.val; if (!er) er = row_split(P, row, cells); }
        if (!er) er = eval_conds(P, row, cells, &keep);
        if (er) { report(bases[b] + e, er); continue; }
        }
        if (!keep) continue;
        live_rows++;
        if (err != ~0ull) continue;  // the host discards rows at/after the first failing entry
        if (P.mode == PM_SCAN) {
          // same split as scan_kernel<PM_SCAN>: fast rows feed integer outputs by stored position, the rest by cell_value
          auto put = [&](int k) {
            Value v;
            int e2 = output_value(P, row, cells, k, &v);  // (the kernel splits this by instantiation: PM_SCAN / PM_PROJ)
            if (e2) { report(bases[b] + e, e2); v.null = true; v.bits = 0; }
            R->data[k].push_back(v.null ? 0 : v.bits); R->nonnull[k].push_back(!v.null);
          };
          if (row.fast) {
            uint32_t prev = 0;
            for (int h = 0; h < 8 && h < P.fast_n; ++h) {
              uint32_t end = fast_end(row, h);
              if (P.fast_out[h] >= 0) { R->data[P.fast_out[h]].push_back(row.fast == 2 ? fast_cell_dyn(row, (uint32_t)h, false, true) : fast_int_cell(row, prev, end, (P.fast_uns >> h) & 1u)); R->nonnull[P.fast_out[h]].push_back(true); }
              prev = end;
            }
            for (int j = 0; j < P.n_out_slow; ++j) put(P.out_slow[j]);
          } else {
            for (int k = 0; k < P.n_out; ++k) put(k);
          }
          R->n_rows++;
        } else if (P.mode == PM_TOPN) {
          TopRow tr;
          int e2 = make_item(P, row, cells, bases[b] + e, &tr.it);
          if (e2) { report(bases[b] + e, e2); continue; }
          for (int k = 0; k < P.n_out; ++k) {
            Value v;
            int e3 = cell_value(P, row, cells, P.out_cols[k], &v);
            if (e3) { v.null = true; v.bits = 0; tr.it.slot = (unsigned)e3; }  // decode errors only matter for surviving rows
            tr.bits.push_back(v.null ? 0 : v.bits); tr.nn.push_back(!v.null);
          }
          top_rows.push_back(std::move(tr));
        } else if (P.mode == PM_AGG) {
          GroupAcc* acc = &single;
          if (P.n_group > 1) {  // scan_body<PM_AGGM>: composite key, no -0.0 folding
            GKey key;
            uint64_t nm = 0;
            int e2 = 0;
            for (int q = 0; q < P.n_group && !e2; ++q) {
              Value v;
              e2 = eval_expr(P, P.groups[q], row, cells, &v, nullptr);
              key.push_back(v.null ? 0ull : v.bits);
              if (v.null) nm |= 1ull << q;
            }
            if (e2) { report(bases[b] + e, e2); continue; }
            key.push_back(nm);
            auto it = groups.find(key);
            if (it == groups.end()) { GroupAcc z; memset(&z, 0, sizeof(z)); it = groups.emplace(key, z).first; group_order.push_back(key); }
            acc = &it->second;
          } else if (P.has_group) {
            Value gk;
            int e2 = eval_expr(P, P.group, row, cells, &gk, nullptr);
            if (e2) { report(bases[b] + e, e2); continue; }
            if (P.group_et == 1 && !gk.null && bits_f64(gk.bits) == 0.0) gk.bits = 0;
            GKey key = {(uint64_t)gk.null, gk.null ? 0ull : gk.bits};
            auto it = groups.find(key);
            if (it == groups.end()) { GroupAcc z; memset(&z, 0, sizeof(z)); it = groups.emplace(key, z).first; group_order.push_back(key); }
            acc = &it->second;
          }
          for (int a = 0; a < P.n_aggs; ++a) {
            const DevAgg g = P.aggs[a];
            Value v;
            int e2 = eval_expr(P, g.arg, row, cells, &v, nullptr);
            if (e2) { report(bases[b] + e, e2); continue; }
            if (v.null) continue;
            acc->w[g.acc_off] += 1;
            if (g.kind == 0) continue;
            if (g.kind >= 3) { uint64_t key = extremum_key(v.bits, g.arg_et, g.arg_unsigned, g.kind == 4); if (key > acc->w[g.acc_off + 1]) acc->w[g.acc_off + 1] = key; continue; }
            if (g.arg_et == 1) f64_acc_add(v.bits, [&](uint32_t d, int64_t x) { acc->w[g.acc_off + 1 + d] += (uint64_t)x; });
            else {
              acc->w[g.acc_off + 1] += v.bits & 0xffffffffull;
              acc->w[g.acc_off + 2] += g.arg_unsigned ? (v.bits >> 32) : (uint64_t)((int64_t)v.bits >> 32);
            }
          }
        }
      }
    }
  }
  if (P.mode == PM_SCAN && cp.scan_limit != ~0ull && R->n_rows >= cp.scan_limit) {
    // BatchLimitExecutor: the first `limit` rows; a failing row beyond them is never reached (limit_executor.rs:55-80)
    for (auto& c : R->data) c.resize((size_t)cp.scan_limit);
    for (auto& c : R->nonnull) c.resize((size_t)cp.scan_limit);
    R->n_rows = cp.scan_limit;
    err = ~0ull;
  }
  if (err != ~0ull) {
    R->dev_err = (int)(err & 0xff); R->err_entry = err >> 8;
    R->status = R->dev_err == DE_IDX_NEW_LAYOUT ? B2_ERR_UNSUPPORTED : (R->dev_err >= 20 && R->dev_err < 30) ? B2_ERR_EVALUATE : (R->dev_err <= 5 && R->dev_err != DE_WRITE_CONFLICT ? B2_ERR_STORAGE : (R->dev_err == DE_WRITE_CONFLICT ? B2_ERR_WRITE_CONFLICT : B2_ERR_CORRUPTED));
    if (P.mode == PM_SCAN) {
      // keep only rows before the failing entry: they were pushed in order, rows after it were skipped above
    }
  }
  if (P.mode == PM_TOPN && err == ~0ull) {
    std::sort(top_rows.begin(), top_rows.end(), [&](const TopRow& a, const TopRow& b) { return item_less(a.it, b.it, P); });
    size_t n = std::min<size_t>(top_rows.size(), P.limit);
    for (size_t i = 0; i < cp.output_offsets.size(); ++i) {
      uint32_t k = cp.output_offsets[i];
      R->kinds[i] = cp.schema[k].kind;
      for (size_t r = 0; r < n; ++r) {
        if (top_rows[r].it.slot && R->status == 0) { R->status = B2_ERR_CORRUPTED; R->dev_err = (int)top_rows[r].it.slot; }
        R->data[i].push_back(top_rows[r].bits[k]); R->nonnull[i].push_back(top_rows[r].nn[k]);
      }
    }
    R->n_rows = R->status ? 0 : n;
  }
  if (P.mode == PM_SCAN) {
    for (int k = 0; k < P.n_out; ++k) R->kinds[k] = cp.schema[cp.output_offsets[k]].kind;
  } else if (P.mode == PM_AGG && err == ~0ull) {
    // agg_result_kernel restated: accumulators -> [aggregates..., group key]
    std::vector<std::pair<GKey, GroupAcc>> gs;
    if (P.has_group) for (auto& k : group_order) gs.push_back({k, groups[k]});
    else if (live_rows > 0) gs.push_back({GKey{0, 0ull}, single});
    size_t ncol = cp.schema.size();
    std::vector<std::vector<uint64_t>> data(ncol);
    std::vector<std::vector<b2_decimal>> dec(ncol);
    std::vector<std::vector<uint8_t>> nn(ncol);
    for (auto& g : gs) {
      int c = 0;
      const uint64_t* acc = g.second.w;
      for (int a = 0; a < P.n_aggs; ++a) {
        const DevAgg ag = P.aggs[a];
        uint64_t cnt = acc[ag.acc_off];
        if (ag.kind == 0 || ag.kind == 2) { data[c].push_back(cnt); nn[c].push_back(1); dec[c].push_back(b2_decimal{}); ++c; }
        if (ag.kind == 1 || ag.kind == 2) {
          bool has = cnt != 0;
          b2_decimal d{};
          if (ag.arg_et == 1) data[c].push_back(has ? f64_acc_round(acc + ag.acc_off + 1) : 0);
          else { if (has) limbs_to_decimal(acc[ag.acc_off + 1], acc[ag.acc_off + 2], ag.arg_unsigned, &d); data[c].push_back(0); }
          dec[c].push_back(d); nn[c].push_back(has);
          ++c;
        }
        if (ag.kind == 3 || ag.kind == 4) {
          bool has = cnt != 0;
          data[c].push_back(has ? extremum_value(acc[ag.acc_off + 1], ag.arg_et, ag.arg_unsigned, ag.kind == 4) : 0);
          dec[c].push_back(b2_decimal{}); nn[c].push_back(has);
          ++c;
        }
      }
      if (P.n_group > 1) {
        for (int q = 0; q < P.n_group; ++q, ++c) {
          bool isnull = (g.first[P.n_group] >> q) & 1;
          data[c].push_back(isnull ? 0 : g.first[q]); nn[c].push_back(!isnull); dec[c].push_back(b2_decimal{});
        }
      } else if (P.has_group) { data[c].push_back(g.first[0] ? 0 : g.first[1]); nn[c].push_back(!g.first[0]); dec[c].push_back(b2_decimal{}); }
    }
    for (size_t i = 0; i < cp.output_offsets.size(); ++i) {
      uint32_t k = cp.output_offsets[i];
      R->kinds[i] = cp.schema[k].kind; R->data[i] = data[k]; R->dec[i] = dec[k]; R->nonnull[i] = nn[k];
    }
    R->n_rows = gs.size();
  }
  return R;
}

// checksum_kernel restated on the host with the same helpers
emu_result* emu_checksum(const b2_key_range* ranges, uint32_t n_ranges, const uint8_t* old_prefix, uint32_t old_len, const uint8_t* new_prefix, uint32_t new_len,
                         const b2_region_source* src) {
  emu_result* R = new emu_result();
  std::vector<BlockView> dviews;
  for (uint32_t i = 0; src->dflt && i < src->n_dflt; ++i) dviews.push_back(view_of(src->dflt[i]));
  DefaultCf dflt; dflt.blocks = dviews.data(); dflt.n_blocks = (uint32_t)dviews.size();
  uint64_t st = ~0ull;
  for (uint32_t i = 0; i < old_len; ++i) st = crc64_table_entry((uint8_t)(st ^ old_prefix[i])) ^ (st >> 8);
  for (uint32_t r = 0; r < n_ranges; ++r) {
    std::vector<uint8_t> lo = encode_memcomparable(ranges[r].start, ranges[r].start_len), hi = encode_memcomparable(ranges[r].end, ranges[r].end_len);
    for (uint32_t b = 0; b < src->n_write; ++b) {
      BlockView blk = view_of(src->write[b]);
      uint32_t e_lo = lower_bound_block(src->write[b], lo), e_hi = lower_bound_block(src->write[b], hi);
      // fast_kernel<PM_CHECKSUM>: clean entries are folded through the linear form of the CRC (b2_device.h): the handle
      // takes one 8-byte table step from the state after the unit's common key bytes, values are XOR-ed right-aligned
      const bool unit_fast = g_emu_fast_front && unit_prefix_ok(blk, e_lo, e_hi) && new_len <= 11;
      uint64_t s11 = st;
      bool prefix_ok = true;
      if (unit_fast) {
        const uint8_t* k0 = blk.kptr(e_lo);
        for (uint32_t j = 0; j < new_len; ++j) prefix_ok = prefix_ok && raw_at(k0, j) == new_prefix[j];
        for (uint32_t j = new_len; j < 11; ++j) s11 = crc64_table_entry((uint8_t)(s11 ^ raw_at(k0, j))) ^ (s11 >> 8);
      }
      std::vector<unsigned long long> T(8 * 256);
      for (uint32_t i = 0; i < 256; ++i) T[i] = crc64_table_entry(i);
      for (uint32_t i = 0; i < 256; ++i) { unsigned long long t = T[i]; for (int kk = 1; kk < 8; ++kk) { t = T[(uint32_t)t & 0xffu] ^ (t >> 8); T[kk * 256 + i] = t; } }
      uint64_t kacc[256] = {0}, vacc[32] = {0}, fast_cnt = 0;
      std::vector<uint32_t> todo;
      if (unit_fast && prefix_ok) {
        for (uint32_t e0 = e_lo; e0 < e_hi; e0 += 32) {
          LaneOut lo[32];
          fast_front_warp(blk, e0, e_lo, e_hi, src->read_ts, src->isolation_level, lo);
          for (uint32_t l = 0; l < 32 && e0 + l < e_hi; ++l) {
            const uint32_t e = e0 + l;
            if (lo[l].flags & FA_PUSH) { todo.push_back(e - ((lo[l].flags & FA_COMMIT) ? 0 : lo[l].push_back)); continue; }
            if (!(lo[l].flags & FA_COMMIT)) continue;
            g_emu_fast_hits++;
            const uint32_t roff = lo[l].row_off, rlen = lo[l].row_len;
            kacc[rlen] ^= crc_step8(T.data(), s11, key_tail_handle_le(lo[l].tail.a, lo[l].tail.b));
            const uint8_t* vend = blk.vptr(e) + roff + rlen;
            for (uint32_t j = 0; 8 * j < rlen; ++j) {  // word j = the 8 value bytes ending 8j bytes before the end
              uint64_t w = 0;
              for (uint32_t q = 0; q < 8; ++q) { uint32_t back = 8 * j + (8 - q); if (back <= rlen) w |= (uint64_t)vend[-(int)back] << (8 * q); }
              vacc[j] ^= w;
            }
            ++fast_cnt;
            R->total_kvs++; R->total_bytes += 19ull + rlen + old_len - new_len;
          }
        }
      } else {
        for (uint32_t e = e_lo; e < e_hi; ++e) todo.push_back(e);
      }
      for (uint32_t e : todo) {
        if (!((e == e_lo) || !same_user_key(blk, e - 1, e))) continue;
        RunOut ro;
        resolve_run(blk, e, e_hi, e_hi, src->read_ts, src->isolation_level, dflt, &ro);
        if (ro.err) { R->status = B2_ERR_STORAGE; R->dev_err = ro.err; continue; }
        if (!ro.found) continue;
        const uint8_t* ek = blk.keys + blk.koff[e];
        uint32_t ekl = blk.koff[e + 1] - blk.koff[e] - 8;
        int rawlen = raw_key_len(ek, ekl);
        if (rawlen < 0) { R->status = B2_ERR_STORAGE; continue; }
        bool ok = (uint32_t)rawlen >= new_len;
        for (uint32_t j = 0; ok && j < new_len; ++j) ok = raw_at(ek, j) == new_prefix[j];
        if (!ok) { R->status = B2_ERR_STORAGE; R->msg = "Wrong prefix expect"; continue; }
        uint64_t c = st;
        for (uint32_t j = new_len; j < (uint32_t)rawlen; ++j) c = crc64_table_entry((uint8_t)(c ^ raw_at(ek, j))) ^ (c >> 8);
        for (uint32_t j = 0; j < ro.val_len; ++j) c = crc64_table_entry((uint8_t)(c ^ ro.val[j])) ^ (c >> 8);
        R->checksum ^= ~c; R->total_kvs++; R->total_bytes += (uint64_t)rawlen + ro.val_len + old_len - new_len;
      }
      // combine the unit's folded state: (odd count ? ~0 : 0) ^ XOR_vlen A^vlen(kacc[vlen]) ^ Lin(vacc)
      uint64_t fold = (fast_cnt & 1) ? ~0ull : 0ull;
      for (uint32_t v = 0; v < 256; ++v) if (kacc[v]) fold ^= crc_advance_zeros(T.data(), kacc[v], v);
      uint64_t lin = 0;
      for (int j = 31; j >= 0; --j) lin = crc_step8(T.data(), lin, vacc[j]);
      R->checksum ^= fold ^ lin;
    }
  }
  return R;
}

void emu_set_fast_front(int on) { g_emu_fast_front = on != 0; }
uint64_t emu_fast_hits(void) { return g_emu_fast_hits; }
int emu_status(emu_result* r) { return r->status; }
int emu_dev_err(emu_result* r) { return r->dev_err; }
uint64_t emu_err_entry(emu_result* r) { return r->err_entry; }
const char* emu_message(emu_result* r) { return r->msg.c_str(); }
uint64_t emu_rows(emu_result* r) { return r->n_rows; }
uint32_t emu_cols(emu_result* r) { return (uint32_t)r->data.size(); }
int emu_col_kind(emu_result* r, uint32_t c) { return r->kinds[c]; }
const uint64_t* emu_col_data(emu_result* r, uint32_t c) { return r->data[c].data(); }
const b2_decimal* emu_col_dec(emu_result* r, uint32_t c) { return r->dec[c].data(); }
const uint8_t* emu_col_nonnull(emu_result* r, uint32_t c) { return r->nonnull[c].data(); }
void emu_stats(emu_result* r, uint64_t* out7) {
  out7[0] = r->processed_keys; out7[1] = r->processed_size; out7[2] = r->met_newer; out7[3] = r->dflt;
  out7[4] = r->checksum; out7[5] = r->total_kvs; out7[6] = r->total_bytes;
}
int emu_parse_decimal(const uint8_t* p, uint32_t n, b2_decimal* out) { return raw_decimal_parse(p, n, out) ? 1 : 0; }
void emu_free(emu_result* r) { delete r; }
int emu_check_supported(const b2_dag_plan* plan, char* msg, size_t cap) {
  CompiledPlan cp; std::string m;
  int rc = compile_plan(plan, &cp, &m);
  snprintf(msg, cap, "%s", m.c_str());
  return rc;
}

// codec hooks: the word-wise varint readers of b2_device.h against the oracle's byte-wise restatement
uint32_t emu_dec_var_u64(const uint8_t* p, uint32_t n, uint64_t* out) { return dec_var_u64(p, n, out); }
uint32_t emu_first_var_int_len(const uint8_t* p, uint32_t n) { return first_var_int_len(p, n); }
uint32_t emu_split_datum(const uint8_t* p, uint32_t n) { int err = 0; uint32_t r = split_datum(p, n, &err); return err ? 0 : r; }

}  // extern "C"
