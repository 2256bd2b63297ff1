// TEST INFRASTRUCTURE — CPU oracle, not product code.
//
// C entry points of the oracle: a CPU restatement of TiKV's coprocessor hot path used (a) as the
// parity checker by tests/ and __graft_entry__.smoke(), (b) as the timed "port" CPU baseline by
// bench.py.  Nothing under tikv_b200/ may link or call this file.
//
// Pinned against the reference's own golden vectors (tests/test_oracle_golden.py): memcomparable
// bytes (tikv_util/src/codec/bytes.rs:352-), Key+ts (txn_types/src/types.rs:890-914), row v2 byte
// arrays (row/v2/encoder_for_test.rs:543-609), write-record cases (txn_types/src/write.rs:504-549),
// table-scan fixture (table_scan_executor.rs:496-512), CRC-64/XZ check value.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <thread>

#include "orc_exec.h"
#include "orc_encode.h"
#include "orc_chunk.h"
#include "orc_gen.h"
#include "orc_sst.h"

using namespace orc;

struct orc_result {
  Error err;
  uint64_t n_rows = 0;
  std::vector<LazyColumn> cols;       // decoded; ET_DECIMAL cols index into decimals
  std::vector<Decimal> decimals;
  std::vector<std::vector<Decimal>> dec_cols;  // per column decimal cells
  std::vector<RawChunkCol> raw_cols;           // per column: cells of a column that stayed Raw (bytes, time, duration, decimal, json)
  std::vector<FieldType> schema;
  Statistics stats;
  int met_newer = NEWER_UNKNOWN;
  uint64_t scanned_rows = 0;
  std::string dec_str;
  Bytes enc_default;  // Chunk.rows_data of every batch, EncodeType::TypeDefault, concatenated (runner.rs:1062-1071)
  Bytes enc_chunk;    // the whole result as one EncodeType::TypeChunk chunk (runner.rs:1072-1085)
  uint64_t warning_cnt = 0;  // SelectResponse.warning_count (runner.rs:845)
};

static bool build_executors(const b2_dag_plan* plan, const b2_key_range* ranges, uint32_t n_ranges, const b2_region_source* src,
                            CfView* w, CfView* l, CfView* d, std::unique_ptr<Executor>* out, TableScanExecutor** scan_out, Error* err) {
  if (plan->n_executors == 0 || (plan->executors[0].tp != B2_EXEC_TABLE_SCAN && plan->executors[0].tp != B2_EXEC_INDEX_SCAN)) {
    *err = Error::make(B2_ERR_UNSUPPORTED, "first executor must be TableScan or IndexScan");
    return false;
  }
  auto scan = std::make_unique<TableScanExecutor>();
  scan->init(plan->executors[0]);
  w->init(src->write, src->n_write);
  d->init(src->dflt, src->dflt ? src->n_dflt : 0);
  l->init(src->lock, src->lock ? 1 : 0);
  RangesScanner& rs = scan->rs;
  rs.w = w; rs.l = l; rs.d = d;
  rs.base_cfg.ts = src->read_ts;
  rs.base_cfg.isolation_level = src->isolation_level;
  rs.base_cfg.check_has_newer_ts_data = src->check_has_newer_ts_data;
  rs.base_cfg.bypass_locks.assign(src->bypass_locks, src->bypass_locks + src->n_bypass_locks);
  rs.base_cfg.access_locks.assign(src->access_locks, src->access_locks + src->n_access_locks);
  for (uint32_t i = 0; i < n_ranges; ++i)
    rs.ranges.emplace_back(Bytes(ranges[i].start, ranges[i].start + ranges[i].start_len), Bytes(ranges[i].end, ranges[i].end + ranges[i].end_len));
  if (plan->executors[0].desc) {  // scan_executor.rs:89-101: ranges in reverse order, each scanned backward
    rs.desc = true;
    std::reverse(rs.ranges.begin(), rs.ranges.end());
  }
  *scan_out = scan.get();
  std::unique_ptr<Executor> cur = std::move(scan);
  for (uint32_t i = 1; i < plan->n_executors; ++i) {
    const b2_executor_desc& e = plan->executors[i];
    if (e.tp == B2_EXEC_SELECTION) {
      auto s = std::make_unique<SelectionExecutor>();
      s->conditions.assign(e.conditions, e.conditions + e.n_conditions);
      s->src = std::move(cur);
      cur = std::move(s);
    } else if (e.tp == B2_EXEC_AGGREGATION || e.tp == B2_EXEC_STREAM_AGG) {
      auto a = std::make_unique<AggExecutor>();
      const auto& sch = cur->schema();
      auto ret_type = [&](const b2_rpn_expr& x, FieldType* ft) {
        const b2_rpn_node& last = x.nodes[x.n_nodes - 1];
        if (last.kind == B2_RPN_COLUMN_REF) *ft = sch[(size_t)last.i64]; else { ft->tp = last.field_tp; ft->flag = last.field_flag; }
      };
      for (uint32_t k = 0; k < e.n_aggrs; ++k) {
        AggFn f; f.kind = e.aggrs[k].kind; f.arg = e.aggrs[k].arg;
        FieldType ft; ret_type(f.arg, &ft);
        const b2_rpn_node& last = f.arg.nodes[f.arg.n_nodes - 1];
        f.arg_et = last.kind == B2_RPN_CONST_REAL ? ET_REAL : (last.kind == B2_RPN_CONST_INT || last.kind == B2_RPN_CONST_UINT ? ET_INT : eval_type_of(ft.tp));
        f.arg_unsigned = ft.is_unsigned();
        if (f.arg_et != ET_INT && f.arg_et != ET_REAL) { *err = Error::make(B2_ERR_UNSUPPORTED, "aggregate over non Int/Real"); return false; }
        if (f.kind != B2_AGG_COUNT && f.kind != B2_AGG_SUM && f.kind != B2_AGG_AVG && f.kind != B2_AGG_MAX && f.kind != B2_AGG_MIN) { *err = Error::make(B2_ERR_UNSUPPORTED, "aggregate kind"); return false; }
        a->fns.push_back(f);
        FieldType cnt; cnt.tp = B2_TP_LONGLONG; cnt.flag = B2_FLAG_UNSIGNED;  // impl_count.rs:35-40
        FieldType sum; if (f.arg_et == ET_REAL) { sum.tp = B2_TP_DOUBLE; } else { sum.tp = B2_TP_NEWDECIMAL; }
        if (f.kind == B2_AGG_COUNT || f.kind == B2_AGG_AVG) a->schema_.push_back(cnt);
        if (f.kind == B2_AGG_SUM || f.kind == B2_AGG_AVG) a->schema_.push_back(sum);
        if (f.kind == B2_AGG_MAX || f.kind == B2_AGG_MIN) a->schema_.push_back(ft);  // the argument's own type, impl_max_min.rs:78-84
      }
      if (e.n_group_by > 1) {
        for (uint32_t q = 0; q < e.n_group_by; ++q) {
          FieldType ft; ret_type(e.group_by[q], &ft);
          EvalType et = eval_type_of(ft.tp);
          if (et != ET_INT && et != ET_REAL) { *err = Error::make(B2_ERR_UNSUPPORTED, "group by non Int/Real"); return false; }
          a->multi_by.push_back(e.group_by[q]); a->multi_ft.push_back(ft); a->schema_.push_back(ft);
        }
      } else if (e.n_group_by == 1) {
        a->has_group = true; a->group_by = e.group_by[0];
        ret_type(a->group_by, &a->group_ft);
        a->group_et = eval_type_of(a->group_ft.tp);
        if (a->group_et != ET_INT && a->group_et != ET_REAL) { *err = Error::make(B2_ERR_UNSUPPORTED, "group by non Int/Real"); return false; }
        a->schema_.push_back(a->group_ft);
      }
      a->src = std::move(cur);
      cur = std::move(a);
    } else if (e.tp == B2_EXEC_TOPN) {
      auto t = std::make_unique<TopNExecutor>();
      for (uint32_t k = 0; k < e.n_order_by; ++k) t->order.push_back({e.order_by[k].expr, e.order_by[k].desc != 0});
      t->n = e.limit;
      t->src = std::move(cur);
      cur = std::move(t);
    } else if (e.tp == B2_EXEC_PROJECTION) {
      auto pj = std::make_unique<ProjectionExecutor>();
      const auto& sch = cur->schema();
      for (uint32_t k = 0; k < e.n_conditions; ++k) {
        pj->exprs.push_back(e.conditions[k]);
        const b2_rpn_node& last = e.conditions[k].nodes[e.conditions[k].n_nodes - 1];
        FieldType ft;
        if (last.kind == B2_RPN_COLUMN_REF) ft = sch[(size_t)last.i64]; else { ft.tp = last.field_tp; ft.flag = last.field_flag; }
        pj->schema_.push_back(ft);
      }
      pj->src = std::move(cur);
      cur = std::move(pj);
    } else if (e.tp == B2_EXEC_LIMIT) {
      auto lm = std::make_unique<LimitExecutor>();
      lm->remaining_rows = (size_t)e.limit;
      lm->is_src_scan_executor = i == 1;  // runner.rs: the child is the scan executor itself
      lm->src = std::move(cur);
      cur = std::move(lm);
    } else { *err = Error::make(B2_ERR_UNSUPPORTED, "executor type " + std::to_string(e.tp)); return false; }
  }
  *out = std::move(cur);
  return true;
}

struct orc_scan { std::vector<Bytes> keys, vals; Statistics st; Error err; int met_newer; };
template <class Scanner>
static orc_scan* mvcc_scan_with(const b2_region_source* src, const uint8_t* lower, size_t lower_len, const uint8_t* upper, size_t upper_len) {
  orc_scan* r = new orc_scan();
  CfView w, l, d;
  w.init(src->write, src->n_write);
  d.init(src->dflt, src->dflt ? src->n_dflt : 0);
  l.init(src->lock, src->lock ? 1 : 0);
  ScannerConfig cfg;
  cfg.ts = src->read_ts; cfg.isolation_level = src->isolation_level; cfg.check_has_newer_ts_data = src->check_has_newer_ts_data;
  cfg.bypass_locks.assign(src->bypass_locks, src->bypass_locks + src->n_bypass_locks);
  cfg.access_locks.assign(src->access_locks, src->access_locks + src->n_access_locks);
  if (lower) { cfg.has_lower = true; cfg.lower_bound.assign(lower, lower + lower_len); }
  if (upper) { cfg.has_upper = true; cfg.upper_bound.assign(upper, upper + upper_len); }
  Scanner fs;
  fs.init(cfg, &w, &l, &d);
  for (;;) {
    ScanOutput so;
    int rc = fs.read_next(&so, &r->err);
    if (rc <= 0) break;
    r->keys.push_back(so.user_key); r->vals.push_back(so.value);
  }
  r->st = fs.statistics; r->met_newer = fs.met_newer_ts_data;
  return r;
}

extern "C" {

// BatchExecutorsRunner::handle_request (runner.rs:739-851): run to drain, collect decoded output rows.
// keep_rows = 0: every batch is still decoded and appended to the response buffers, but the buffers are recycled per
// batch (the reference streams / pages its response chunks, runner.rs:790-806) so a timed run does not hold the
// whole result in memory.
static int orc_dag_handle_impl(const b2_dag_plan* plan, const b2_key_range* ranges, uint32_t n_ranges, const b2_region_source* src, orc_result** out, int keep_rows) {
  orc_result* res = new orc_result();
  *out = res;
  warning_count() = 0;
  struct KeepWarnings { orc_result* r; ~KeepWarnings() { r->warning_cnt = warning_count(); } } keep_warnings{res};
  CfView w, l, d;
  std::unique_ptr<Executor> root;
  TableScanExecutor* scan = nullptr;
  if (!build_executors(plan, ranges, n_ranges, src, &w, &l, &d, &root, &scan, &res->err)) return res->err.status;
  const auto& sch = root->schema();
  std::vector<uint32_t> offs;
  if (plan->output_offsets) offs.assign(plan->output_offsets, plan->output_offsets + plan->n_output_offsets);
  else for (uint32_t i = 0; i < sch.size(); ++i) offs.push_back(i);
  res->cols.assign(offs.size(), LazyColumn());
  res->dec_cols.assign(offs.size(), std::vector<Decimal>());
  res->raw_cols.assign(offs.size(), RawChunkCol());
  for (size_t k = 0; k < offs.size(); ++k) { res->schema.push_back(sch[offs[k]]); res->cols[k].decoded = true; res->cols[k].et = eval_type_of(sch[offs[k]].tp); }
  size_t batch_size = BATCH_INITIAL_SIZE;
  AggExecutor* agg = dynamic_cast<AggExecutor*>(root.get());
  for (;;) {
    Batch b;
    root->next_batch(batch_size, &b);
    // encode_result_to_chunk, TypeDefault arm: row by row, output offset by output offset, on the batch as the
    // executors left it (a column no expression touched is still Raw and goes out as the stored datum bytes)
    if (keep_rows && !b.cols.empty())
      for (size_t r : b.logical_rows)
        for (size_t k = 0; k < offs.size(); ++k) encode_cell_default(res->enc_default, b.cols[offs[k]], r, sch[offs[k]], agg ? &agg->dec_col : nullptr);
    for (size_t k = 0; k < offs.size(); ++k) {
      LazyColumn& c = b.cols.empty() ? res->cols[k] : b.cols[offs[k]];
      if (b.cols.empty()) break;
      std::string perr;
      // a column the executors left Raw whose type they never decode goes straight to its chunk cells (Column::from_raw_datums)
      if (!c.decoded && raw_kind_of(sch[offs[k]].tp) != RK_NONE) {
        RawChunkCol& rc = res->raw_cols[k];
        rc.kind = raw_kind_of(sch[offs[k]].tp);
        bool ok = true;
        for (size_t r : b.logical_rows) if (!append_raw_datum(rc, c.raw_get(r), sch[offs[k]], &perr)) { ok = false; break; }
        if (!ok) { if (b.err.ok()) b.err = Error::make(B2_ERR_CORRUPTED, perr); b.logical_rows.clear(); break; }
        continue;
      }
      if (!ensure_decoded(c, sch[offs[k]], b.logical_rows, &perr)) { if (b.err.ok()) b.err = Error::make(B2_ERR_CORRUPTED, perr); b.logical_rows.clear(); break; }
      if (c.et == ET_TIME || c.et == ET_DURATION) {  // a DateTime / Duration column an expression decoded: the same 8-byte chunk cells (column.rs:446-496)
        RawChunkCol& rc = res->raw_cols[k];
        rc.kind = c.et == ET_TIME ? RK_TIME : RK_DURATION;
        for (size_t r : b.logical_rows) { rc.nn.push_back(c.nn[r]); put_le64(rc.data, c.nn[r] ? (uint64_t)c.i64[r] : 0); }
      }
    }
    if (!b.cols.empty())
      for (size_t r : b.logical_rows) {
        for (size_t k = 0; k < offs.size(); ++k) {
          const LazyColumn& c = b.cols[offs[k]];
          LazyColumn& o = res->cols[k];
          if (res->raw_cols[k].kind != RK_NONE) continue;
          o.nn.push_back(c.nn[r]);
          if (c.et == ET_REAL) o.f64.push_back(c.f64[r]);
          else if (c.et == ET_DECIMAL) { res->dec_cols[k].push_back(agg->dec_col[(size_t)c.i64[r]]); o.i64.push_back(0); }
          else o.i64.push_back(c.i64[r]);
        }
        res->n_rows++;
      }
    if (!keep_rows)
      for (size_t k = 0; k < offs.size(); ++k) { res->cols[k].nn.clear(); res->cols[k].i64.clear(); res->cols[k].f64.clear(); res->dec_cols[k].clear(); res->raw_cols[k].clear(); }
    if (!b.err.ok()) { res->err = b.err; break; }
    if (b.is_drained) break;
    if (batch_size < BATCH_MAX_SIZE) { batch_size *= BATCH_GROW_FACTOR; if (batch_size > BATCH_MAX_SIZE) batch_size = BATCH_MAX_SIZE; }  // runner.rs:1098-1105
  }
  scan->rs.accumulate();
  res->stats = scan->rs.total;
  res->met_newer = scan->rs.met_newer;
  res->scanned_rows = scan->rs.rows;
  if (keep_rows)  // TypeChunk arm over the rows produced (chunk boundaries are the server's choice: one chunk here)
    for (size_t k = 0; k < offs.size(); ++k) {
      const LazyColumn& c = res->cols[k];
      if (res->raw_cols[k].kind != RK_NONE) { encode_raw_chunk(res->enc_chunk, res->raw_cols[k]); continue; }
      encode_column_chunk(res->enc_chunk, res->schema[k].tp, c.nn, c.et == ET_REAL ? nullptr : c.i64.data(), c.et == ET_REAL ? c.f64.data() : nullptr,
                          c.et == ET_DECIMAL ? res->dec_cols[k].data() : nullptr);
    }
  return res->err.status;
}

int orc_dag_handle(const b2_dag_plan* plan, const b2_key_range* ranges, uint32_t n_ranges, const b2_region_source* src, orc_result** out) {
  return orc_dag_handle_impl(plan, ranges, n_ranges, src, out, 1);
}

uint64_t orc_result_rows(orc_result* r) { return r->n_rows; }
uint32_t orc_result_cols(orc_result* r) { return (uint32_t)r->cols.size(); }
int orc_result_col_kind(orc_result* r, uint32_t c) {
  switch (r->raw_cols[c].kind) {
    case RK_BYTES: return B2_COL_BYTES;
    case RK_TIME: return B2_COL_TIME;
    case RK_DURATION: return B2_COL_DURATION;
    case RK_DECIMAL: return B2_COL_DECIMAL;
    case RK_JSON: return B2_COL_JSON;
    default: break;
  }
  return r->cols[c].et == ET_REAL ? B2_COL_F64 : (r->cols[c].et == ET_DECIMAL ? B2_COL_DECIMAL : B2_COL_I64);
}
// cells of a column that stayed Raw: fixed cells (8 / 40 bytes) or byte heap + (rows + 1) offsets
const uint8_t* orc_result_col_raw(orc_result* r, uint32_t c, uint64_t* data_len, const int64_t** offsets) {
  const RawChunkCol& rc = r->raw_cols[c];
  if (rc.kind == RK_NONE) { *data_len = 0; *offsets = nullptr; return nullptr; }
  *data_len = rc.data.size(); *offsets = rc.var() ? rc.offsets.data() : nullptr;
  return rc.data.data();
}
const int64_t* orc_result_col_i64(orc_result* r, uint32_t c) { return r->cols[c].i64.data(); }
const double* orc_result_col_f64(orc_result* r, uint32_t c) { return r->cols[c].f64.data(); }
const uint8_t* orc_result_col_nonnull(orc_result* r, uint32_t c) { return r->raw_cols[c].kind != RK_NONE ? r->raw_cols[c].nn.data() : r->cols[c].nn.data(); }
const b2_decimal* orc_result_col_decimal(orc_result* r, uint32_t c) { return (const b2_decimal*)r->dec_cols[c].data(); }
const char* orc_result_decimal_str(orc_result* r, uint32_t c, uint64_t row) { r->dec_str = dec_to_string(r->dec_cols[c][row]); return r->dec_str.c_str(); }
int orc_result_status(orc_result* r) { return r->err.status; }
int orc_result_mysql_code(orc_result* r) { return r->err.mysql_code; }
const char* orc_result_message(orc_result* r) { return r->err.msg.c_str(); }
void orc_result_stats(orc_result* r, uint64_t* out8) {
  out8[0] = r->stats.write.next; out8[1] = r->stats.write.seek; out8[2] = r->stats.write.over_seek_bound; out8[3] = r->stats.write.processed_keys;
  out8[4] = r->stats.processed_size; out8[5] = r->stats.data.processed_keys; out8[6] = r->stats.lock.processed_keys; out8[7] = (uint64_t)(int64_t)r->met_newer;
}
const uint8_t* orc_result_encoded(orc_result* r, int encode_type, uint64_t* len) {
  const Bytes& b = encode_type == 0 ? r->enc_default : r->enc_chunk;
  *len = b.size();
  return b.data();
}
uint64_t orc_result_warning_count(orc_result* r) { return r->warning_cnt; }
void orc_result_free(orc_result* r) { delete r; }

// ChecksumContext::handle_request (src/coprocessor/checksum.rs:59-98)
int orc_checksum_handle(const b2_key_range* ranges, uint32_t n_ranges, const uint8_t* old_prefix, uint32_t old_len, const uint8_t* new_prefix,
                        uint32_t new_len, const b2_region_source* src, b2_checksum_response* out, char* errbuf, size_t errbuf_len) {
  CfView w, l, d;
  w.init(src->write, src->n_write);
  d.init(src->dflt, src->dflt ? src->n_dflt : 0);
  l.init(src->lock, src->lock ? 1 : 0);
  RangesScanner rs;
  rs.w = &w; rs.l = &l; rs.d = &d;
  rs.base_cfg.ts = src->read_ts;
  rs.base_cfg.isolation_level = src->isolation_level;
  rs.base_cfg.bypass_locks.assign(src->bypass_locks, src->bypass_locks + src->n_bypass_locks);
  for (uint32_t i = 0; i < n_ranges; ++i)
    rs.ranges.emplace_back(Bytes(ranges[i].start, ranges[i].start + ranges[i].start_len), Bytes(ranges[i].end, ranges[i].end + ranges[i].end_len));
  Crc64Digest prefix_digest;
  prefix_digest.write(old_prefix, old_len);
  uint64_t checksum = 0, total_kvs = 0, total_bytes = 0;
  for (;;) {
    Bytes k; ScanOutput so; Error e;
    int r = rs.next(&k, &so, &e);
    if (r < 0) { snprintf(errbuf, errbuf_len, "%s", e.msg.c_str()); return e.status; }
    if (r == 0) break;
    if (k.size() < new_len || memcmp(k.data(), new_prefix, new_len) != 0) { snprintf(errbuf, errbuf_len, "Wrong prefix expect"); return B2_ERR_STORAGE; }
    Crc64Digest dg = prefix_digest;  // checksum_crc64_xor :105-114
    dg.write(k.data() + new_len, k.size() - new_len);
    dg.write(so.value.data(), so.value.size());
    checksum ^= dg.sum64();
    total_kvs += 1;
    total_bytes += k.size() + so.value.size() + old_len - new_len;
  }
  out->checksum = checksum; out->total_kvs = total_kvs; out->total_bytes = total_bytes;
  return B2_OK;
}

// One region task per thread (TiKV read pool shape): run `n_tasks` independent DAG requests, each over its
// own region source, on `n_threads` threads.  Returns total rows produced; used by bench.py's CPU baseline.
uint64_t orc_dag_handle_parallel(const b2_dag_plan* plan, const b2_key_range* ranges, uint32_t n_ranges, const b2_region_source* srcs, uint32_t n_tasks,
                                 uint32_t n_threads, uint64_t* scanned_rows_out, int* status_out) {
  std::vector<uint64_t> rows(n_tasks, 0), scanned(n_tasks, 0);
  std::vector<int> status(n_tasks, 0);
  std::vector<std::thread> th;
  std::atomic<uint32_t> next{0};
  for (uint32_t t = 0; t < n_threads; ++t)
    th.emplace_back([&]() {
      for (;;) {
        uint32_t i = next.fetch_add(1);
        if (i >= n_tasks) break;
        orc_result* r = nullptr;
        status[i] = orc_dag_handle_impl(plan, ranges, n_ranges, &srcs[i], &r, 0);
        rows[i] = r->n_rows; scanned[i] = r->scanned_rows;
        delete r;
      }
    });
  for (auto& x : th) x.join();
  uint64_t total = 0, sc = 0; int st = 0;
  for (uint32_t i = 0; i < n_tasks; ++i) { total += rows[i]; sc += scanned[i]; if (status[i]) st = status[i]; }
  *scanned_rows_out = sc; *status_out = st;
  return total;
}


// ---- bench.py's CPU arm: host-generated regions (orc_gen.h), one region task per thread ----------------------------
// `n_tasks` regions of `rows_per_task` consecutive handles each, starting at spec->first_handle: the same table the device
// generator of the product builds, so a GPU request over the same handles must return the same result.
struct orc_bench {
  std::vector<GenBlock> blocks;
  std::vector<b2_cf_block> views;
  std::vector<b2_region_source> srcs;
};
orc_bench* orc_bench_create(const b2_gen_spec* spec, uint32_t n_tasks, uint64_t rows_per_task, uint64_t read_ts, uint32_t n_threads) {
  orc_bench* h = new orc_bench();
  h->blocks.resize(n_tasks); h->views.resize(n_tasks); h->srcs.resize(n_tasks);
  std::vector<std::thread> th;
  std::atomic<uint32_t> next{0};
  for (uint32_t t = 0; t < std::max(1u, n_threads); ++t)
    th.emplace_back([&]() {
      for (;;) {
        uint32_t i = next.fetch_add(1);
        if (i >= n_tasks) break;
        b2_gen_spec s = *spec;
        s.first_handle = spec->first_handle + (uint64_t)i * rows_per_task;
        s.n_rows = rows_per_task;
        gen_block(s, &h->blocks[i]);
      }
    });
  for (auto& x : th) x.join();
  for (uint32_t i = 0; i < n_tasks; ++i) {
    h->views[i] = h->blocks[i].view();
    b2_region_source& r = h->srcs[i];
    memset(&r, 0, sizeof(r));
    r.location = B2_LOC_HOST; r.write = &h->views[i]; r.n_write = 1; r.read_ts = read_ts; r.isolation_level = B2_ISO_SI; r.check_has_newer_ts_data = 1;
  }
  return h;
}
const b2_region_source* orc_bench_source(orc_bench* h, uint32_t task) { return &h->srcs[task]; }
uint64_t orc_bench_bytes(orc_bench* h) {  // key bytes + value bytes + 8 bytes of offsets per entry
  uint64_t n = 0;
  for (auto& b : h->blocks) n += b.koff.back() + b.voff.back() + 8ull * (b.koff.size() - 1);
  return n;
}
uint64_t orc_bench_step(orc_bench* h, const b2_dag_plan* plan, const b2_key_range* ranges, uint32_t n_ranges, uint32_t n_threads, uint64_t* scanned_rows_out,
                        int* status_out) {
  return orc_dag_handle_parallel(plan, ranges, n_ranges, h->srcs.data(), (uint32_t)h->srcs.size(), n_threads, scanned_rows_out, status_out);
}
int orc_checksum_handle(const b2_key_range* ranges, uint32_t n_ranges, const uint8_t* old_prefix, uint32_t old_len, const uint8_t* new_prefix, uint32_t new_len,
                        const b2_region_source* src, b2_checksum_response* out, char* msg, size_t msg_cap);
// checksum.rs:59-98 over every region, one task per thread; XOR / sums of the per-region responses
int orc_bench_checksum_step(orc_bench* h, const b2_key_range* ranges, uint32_t n_ranges, uint32_t n_threads, b2_checksum_response* out) {
  std::vector<b2_checksum_response> res(h->srcs.size());
  std::vector<int> status(h->srcs.size(), 0);
  std::vector<std::thread> th;
  std::atomic<uint32_t> next{0};
  for (uint32_t t = 0; t < std::max(1u, n_threads); ++t)
    th.emplace_back([&]() {
      char msg[64];
      for (;;) {
        uint32_t i = next.fetch_add(1);
        if (i >= h->srcs.size()) break;
        status[i] = orc_checksum_handle(ranges, n_ranges, nullptr, 0, nullptr, 0, &h->srcs[i], &res[i], msg, sizeof(msg));
      }
    });
  for (auto& x : th) x.join();
  memset(out, 0, sizeof(*out));
  int st = 0;
  for (size_t i = 0; i < res.size(); ++i) { out->checksum ^= res[i].checksum; out->total_kvs += res[i].total_kvs; out->total_bytes += res[i].total_bytes; if (status[i]) st = status[i]; }
  return st;
}
void orc_bench_free(orc_bench* h) { delete h; }

// Raw MVCC scan over [lower, upper) (encoded user keys; NULL = unbounded) — pins ForwardScanner against the
// reference's scanner unit tests (forward.rs:1179-1727), including exact next/seek statistics.
orc_scan* orc_mvcc_scan(const b2_region_source* src, const uint8_t* lower, size_t lower_len, const uint8_t* upper, size_t upper_len) {
  return mvcc_scan_with<ForwardScanner>(src, lower, lower_len, upper, upper_len);
}
// the same over BackwardScanner (backward.rs tests :524-1576): keys come back in descending order
orc_scan* orc_mvcc_scan_backward(const b2_region_source* src, const uint8_t* lower, size_t lower_len, const uint8_t* upper, size_t upper_len) {
  return mvcc_scan_with<BackwardScanner>(src, lower, lower_len, upper, upper_len);
}
uint64_t orc_scan_rows(orc_scan* r) { return r->keys.size(); }
const uint8_t* orc_scan_key(orc_scan* r, uint64_t i, size_t* len) { *len = r->keys[i].size(); return r->keys[i].data(); }
const uint8_t* orc_scan_val(orc_scan* r, uint64_t i, size_t* len) { *len = r->vals[i].size(); return r->vals[i].data(); }
int orc_scan_status(orc_scan* r) { return r->err.status; }
void orc_scan_stats(orc_scan* r, uint64_t* out8) {
  out8[0] = r->st.write.next; out8[1] = r->st.write.seek; out8[2] = r->st.write.over_seek_bound; out8[3] = r->st.write.processed_keys;
  out8[4] = r->st.processed_size; out8[5] = r->st.data.processed_keys; out8[6] = r->st.lock.processed_keys; out8[7] = (uint64_t)(int64_t)r->met_newer;
}
void orc_scan_stats_backward(orc_scan* r, uint64_t* out2) { out2[0] = r->st.write.prev; out2[1] = r->st.write.seek_for_prev; }
void orc_scan_free(orc_scan* r) { delete r; }

// ---- codec hooks for the golden-vector tests ----
size_t orc_encode_bytes(const uint8_t* p, size_t n, uint8_t* out) { Bytes b; encode_bytes(b, Slice(p, n)); memcpy(out, b.data(), b.size()); return b.size(); }
int64_t orc_decode_bytes(const uint8_t* p, size_t n, uint8_t* out, size_t* out_len) {
  Bytes k; size_t c = decode_bytes(Slice(p, n), &k);
  if (c == (size_t)-1) return -1;
  memcpy(out, k.data(), k.size()); *out_len = k.size();
  return (int64_t)c;
}
size_t orc_key_append_ts(const uint8_t* p, size_t n, uint64_t ts, uint8_t* out) { Bytes b = key_append_ts(Bytes(p, p + n), ts); memcpy(out, b.data(), b.size()); return b.size(); }
size_t orc_encode_var_u64(uint64_t v, uint8_t* out) { Bytes b; encode_var_u64(b, v); memcpy(out, b.data(), b.size()); return b.size(); }
size_t orc_encode_var_i64(int64_t v, uint8_t* out) { Bytes b; encode_var_i64(b, v); memcpy(out, b.data(), b.size()); return b.size(); }
size_t orc_decode_var_u64(const uint8_t* p, size_t n, uint64_t* v) { return decode_var_u64(Slice(p, n), v); }
size_t orc_decode_var_i64(const uint8_t* p, size_t n, int64_t* v) { return decode_var_i64(Slice(p, n), v); }
uint64_t orc_encode_i64_cmp(int64_t v) { return (uint64_t)v ^ SIGN_MARK; }
uint64_t orc_encode_f64_cmp(double f) { return encode_f64_to_cmp_u64(f); }
double orc_decode_f64_cmp(uint64_t u) { return decode_cmp_u64_to_f64(u); }
size_t orc_encode_row_key(int64_t table_id, int64_t handle, uint8_t* out) { Bytes b = encode_row_key(table_id, handle); memcpy(out, b.data(), b.size()); return b.size(); }
uint64_t orc_crc64(const uint8_t* p, size_t n) { Crc64Digest d; d.write(p, n); return d.sum64(); }
size_t orc_split_datum(const uint8_t* p, size_t n) { std::string e; return split_datum(Slice(p, n), &e); }

// write record: returns 0 ok. fields: [type, start_ts, has_short, short_off, short_len, overlapped, has_gc_fence, gc_fence, lc_kind, lc_ts, lc_versions, txn_source]
int orc_write_parse(const uint8_t* p, size_t n, uint64_t* f12) {
  WriteRef w; std::string e;
  if (!write_parse(Slice(p, n), &w, &e)) return 1;
  f12[0] = w.write_type; f12[1] = w.start_ts; f12[2] = w.has_short_value; f12[3] = w.has_short_value ? (uint64_t)(w.short_value.p - p) : 0;
  f12[4] = w.has_short_value ? w.short_value.n : 0; f12[5] = w.has_overlapped_rollback; f12[6] = w.has_gc_fence; f12[7] = w.gc_fence;
  f12[8] = w.last_change; f12[9] = w.last_change_ts; f12[10] = w.estimated_versions_to_last_change; f12[11] = w.txn_source;
  return 0;
}
int orc_write_check_gc_fence(const uint8_t* p, size_t n, uint64_t read_ts) {
  WriteRef w; std::string e;
  if (!write_parse(Slice(p, n), &w, &e)) return -1;
  return write_check_gc_fence_as_latest_version(w, read_ts);
}

// decimal hooks
void orc_decimal_from_i64(int64_t v, b2_decimal* out) { Decimal d = dec_from_i64(v); memcpy(out, &d, 40); }
void orc_decimal_from_u64(uint64_t v, b2_decimal* out) { Decimal d = dec_from_u64(v); memcpy(out, &d, 40); }
int orc_decimal_add(const b2_decimal* a, const b2_decimal* b, b2_decimal* out) { Decimal r; int st = dec_add(*(const Decimal*)a, *(const Decimal*)b, &r); memcpy(out, &r, 40); return st; }
size_t orc_decimal_to_string(const b2_decimal* a, char* out, size_t cap) { std::string s = dec_to_string(*(const Decimal*)a); snprintf(out, cap, "%s", s.c_str()); return s.size(); }
size_t orc_decimal_write(const b2_decimal* a, int prec, int frac, uint8_t* out) {  // prec < 0: prec_and_frac()
  Bytes b; uint8_t p = (uint8_t)prec, f = (uint8_t)frac;
  if (prec < 0) dec_prec_and_frac(*(const Decimal*)a, &p, &f);
  dec_write(b, *(const Decimal*)a, p, f);
  memcpy(out, b.data(), b.size());
  return b.size();
}
int orc_decimal_cmp(const b2_decimal* a, const b2_decimal* b) { return dec_cmp(*(const Decimal*)a, *(const Decimal*)b); }

// ---- RocksDB data blocks (orc_sst.h): builder and iterator for the block-reader tests ----
struct orc_sst_buf { std::string data, keys, vals; std::vector<uint64_t> offs; std::vector<uint32_t> koff, voff; };
void* orc_sst_build(const b2_cf_block* flat, uint32_t restart_interval, uint32_t block_size, uint32_t entries_per_block, uint32_t key_prefix_len, uint8_t key_prefix_byte,
                    uint32_t key_suffix_len, uint32_t trailer_len) {
  SstOptions o;
  o.restart_interval = restart_interval; o.block_size = block_size; o.entries_per_block = entries_per_block; o.key_prefix_len = key_prefix_len; o.key_prefix_byte = key_prefix_byte;
  o.key_suffix_len = key_suffix_len; o.trailer_len = trailer_len;
  auto* r = new orc_sst_buf();
  sst_build(flat->keys, flat->key_offs, flat->vals, flat->val_offs, flat->n, o, &r->data, &r->offs);
  return r;
}
// 0 ok, 1 corrupted, 2 unsupported; *out holds the flat block (orc_sst_flat)
int orc_sst_decode(const uint8_t* data, const uint64_t* block_offs, uint32_t n_blocks, uint32_t trailer_len, uint32_t key_prefix_len, uint32_t key_suffix_len, void** out) {
  auto* r = new orc_sst_buf();
  int rc = sst_decode(data, block_offs, n_blocks, trailer_len, key_prefix_len, key_suffix_len, &r->keys, &r->koff, &r->vals, &r->voff);
  *out = r;
  return rc;
}
const uint8_t* orc_sst_data(void* h, uint64_t* len, const uint64_t** offs, uint32_t* n_blocks) {
  auto* r = (orc_sst_buf*)h;
  *len = r->data.size(); *offs = r->offs.data(); *n_blocks = (uint32_t)r->offs.size() - 1;
  return (const uint8_t*)r->data.data();
}
void orc_sst_flat(void* h, b2_cf_block* out) {
  auto* r = (orc_sst_buf*)h;
  out->keys = (const uint8_t*)r->keys.data(); out->key_offs = r->koff.data(); out->vals = (const uint8_t*)r->vals.data(); out->val_offs = r->voff.data();
  out->n = (uint32_t)r->koff.size() - 1;
}
void orc_sst_free(void* h) { delete (orc_sst_buf*)h; }

}  // extern "C"
