// TEST INFRASTRUCTURE — CPU oracle, not product code.
// Batch executors restated row-at-a-time / 1024-row batches, same algorithmic shape as the reference:
//   RangesScanner / TikvStorage   tidb_query_common/src/storage/scanner.rs:122-180, src/coprocessor/dag/storage_impl.rs:39-86
//   ScanExecutor                  tidb_query_executors/src/util/scan_executor.rs:114-169, 226-261
//   TableScanExecutorImpl         tidb_query_executors/src/table_scan_executor.rs:56-150, 200-281, 365-475
//   LazyBatchColumn               tidb_query_datatype/src/codec/batch/lazy_column.rs:165-221
//   RpnExpression::eval           tidb_query_expr/src/types/expr_eval.rs:206-330 (+ impl_compare.rs, impl_op.rs, impl_arithmetic.rs)
//   BatchSelectionExecutor        selection_executor.rs:81-195
//   AggregationExecutor           util/aggr_executor.rs:206-303, simple_aggr_executor.rs:139-257,
//                                 fast_hash_aggr_executor.rs:267-457, util/hash_aggr_helper.rs:21-75
//   aggregate states              tidb_query_aggr/src/{impl_count.rs,impl_sum.rs,impl_avg.rs,util.rs:31-66}
//   BatchTopNExecutor             top_n_executor.rs:184-273, util/top_n_heap.rs:36-222, scalar.rs:374-411
//   BatchExecutorsRunner          runner.rs:739-851, 962-1013 (batch growth 32 -> x2 -> 1024)
#pragma once
#include <map>
#include <algorithm>
#include <cmath>
#include <memory>
#include <unordered_map>

#include "orc_decimal.h"
#include "orc_mvcc.h"

namespace orc {

const size_t BATCH_INITIAL_SIZE = 32;  // runner.rs:39
const size_t BATCH_MAX_SIZE = 1024;    // logical_rows.rs:5
const size_t BATCH_GROW_FACTOR = 2;    // runner.rs:51

enum EvalType { ET_INT, ET_REAL, ET_DECIMAL, ET_OTHER, ET_TIME /* DATE / DATETIME: CoreTime bits */, ET_DURATION /* nanoseconds */, ET_BYTES, ET_DEC /* Decimal operand of a comparison */ };
struct FieldType { int tp = 0; uint32_t flag = 0; int decimal = 0; bool is_unsigned() const { return flag & B2_FLAG_UNSIGNED; } };

inline EvalType eval_type_of(int tp) {  // def/eval_type.rs:53-95
  switch (tp) {
    case B2_TP_TINY: case B2_TP_SHORT: case B2_TP_INT24: case B2_TP_LONG: case B2_TP_LONGLONG: case B2_TP_YEAR: case B2_TP_BIT:
      return ET_INT;
    case B2_TP_FLOAT: case B2_TP_DOUBLE: return ET_REAL;
    case B2_TP_NEWDECIMAL: return ET_DECIMAL;
    default: return ET_OTHER;
  }
}

// ---- datum codec (codec/datum.rs:35-49, 1117-1155; datum_codec.rs:401-446) ----
enum { NIL_FLAG = 0, BYTES_FLAG = 1, COMPACT_BYTES_FLAG = 2, INT_FLAG = 3, UINT_FLAG = 4, FLOAT_FLAG = 5, DECIMAL_FLAG = 6,
       DURATION_FLAG = 7, VAR_INT_FLAG = 8, VAR_UINT_FLAG = 9, JSON_FLAG = 10 };

// split_datum (desc = false). returns total datum length (flag + payload) or 0 with err set
inline size_t split_datum(Slice buf, std::string* err) {
  if (buf.empty()) { *err = "datum is too short"; return 0; }
  size_t pos;
  Slice rest = buf.sub(1);
  switch (buf[0]) {
    case INT_FLAG: case UINT_FLAG: case FLOAT_FLAG: case DURATION_FLAG: pos = 8; break;
    case BYTES_FLAG: pos = memcmp_first_encoded_len(rest); break;
    case COMPACT_BYTES_FLAG: {
      int64_t v; size_t n = decode_var_i64(rest, &v);
      if (!n) pos = rest.n; else { size_t r = (size_t)v + n; pos = r < rest.n ? r : rest.n; }
      break;
    }
    case NIL_FLAG: pos = 0; break;
    case DECIMAL_FLAG: {  // mysql::dec_encoded_len decimal.rs:169-190
      if (rest.n < 2) { *err = "decimal too short"; return 0; }
      uint8_t prec = rest[0], frac = rest[1];
      if (prec < frac) { *err = "invalid decimal"; return 0; }
      static const uint8_t DIG_2_BYTES[10] = {0, 1, 1, 2, 2, 3, 3, 4, 4, 4};
      uint8_t int_cnt = prec - frac;
      pos = (int_cnt / 9) * 4 + DIG_2_BYTES[int_cnt % 9] + (frac / 9) * 4 + DIG_2_BYTES[frac % 9] + 2;
      break;
    }
    case VAR_INT_FLAG: case VAR_UINT_FLAG: pos = first_var_int_len(rest); break;
    default: *err = "unsupported data type `" + std::to_string(buf[0]) + "`"; return 0;
  }
  if (buf.n < pos + 1) { *err = "datum is too short"; return 0; }
  return pos + 1;
}

inline bool decode_int_datum(Slice d, bool* is_null, int64_t* out, std::string* err) {  // datum_codec.rs:401-421
  if (d.empty()) { *err = "Failed to decode datum flag"; return false; }
  Slice p = d.sub(1);
  *is_null = false;
  switch (d[0]) {
    case NIL_FLAG: *is_null = true; *out = 0; return true;
    case INT_FLAG: if (p.n < 8) { *err = "unexpected eof"; return false; } *out = decode_i64(p.p); return true;
    case UINT_FLAG: if (p.n < 8) { *err = "unexpected eof"; return false; } *out = (int64_t)get_u64_be(p.p); return true;
    case VAR_INT_FLAG: if (!decode_var_i64(p, out)) { *err = "unexpected eof"; return false; } return true;
    case VAR_UINT_FLAG: { uint64_t u; if (!decode_var_u64(p, &u)) { *err = "unexpected eof"; return false; } *out = (int64_t)u; return true; }
    default: *err = "Unsupported datum flag " + std::to_string(d[0]) + " for Int vector"; return false;
  }
}
inline bool decode_real_datum(Slice d, int tp, bool* is_null, double* out, std::string* err) {  // datum_codec.rs:423-446
  if (d.empty()) { *err = "Failed to decode datum flag"; return false; }
  Slice p = d.sub(1);
  *is_null = false;
  switch (d[0]) {
    case NIL_FLAG: *is_null = true; *out = 0; return true;
    case FLOAT_FLAG: {
      if (p.n < 8) { *err = "unexpected eof"; return false; }
      double v = decode_cmp_u64_to_f64(get_u64_be(p.p));
      if (tp == B2_TP_FLOAT) v = (double)(float)v;
      if (std::isnan(v)) { *is_null = true; v = 0; }
      *out = v;
      return true;
    }
    default: *err = "Unsupported datum flag " + std::to_string(d[0]) + " for Real vector"; return false;
  }
}

// ---- columns ----
struct LazyColumn {
  bool decoded = false;
  EvalType et = ET_OTHER;
  // raw BufferVec (tikv_util/src/buffer_vec.rs:9-13)
  Bytes raw_data; std::vector<size_t> raw_offsets;
  // decoded ChunkedVecSized (chunked_vec_sized.rs:17-22); nn[i] = 1 non-null
  std::vector<int64_t> i64; std::vector<double> f64; std::vector<uint8_t> nn;
  size_t len() const { return decoded ? nn.size() : raw_offsets.size(); }
  void raw_push(Slice s) { raw_offsets.push_back(raw_data.size()); raw_data.insert(raw_data.end(), s.p, s.p + s.n); }
  Slice raw_get(size_t i) const {
    size_t a = raw_offsets[i], b = i + 1 < raw_offsets.size() ? raw_offsets[i + 1] : raw_data.size();
    return Slice(raw_data.data() + a, b - a);
  }
  void raw_truncate(size_t n) { if (n < raw_offsets.size()) { raw_data.resize(raw_offsets[n]); raw_offsets.resize(n); } }
  void push_int(bool non_null, int64_t v) { i64.push_back(non_null ? v : 0); nn.push_back(non_null); }
};

// Time::from_packed_u64 for DATE / DATETIME (TIMESTAMP converts through the session time zone: not restated).
inline bool time_from_packed(uint64_t value, int tp, int decimal, uint64_t* bits, std::string* err) {
  if (tp == B2_TP_TIMESTAMP) { *err = "oracle does not restate TIMESTAMP time-zone conversion"; return false; }
  if (decimal != -1 && (decimal < 0 || decimal > 6)) { *err = "Invalid fsp"; return false; }
  const uint64_t fsp = decimal == -1 ? 0 : (uint64_t)decimal;
  const bool date = tp == B2_TP_DATE;
  const uint64_t fsp_tt = date ? 0xeull : (fsp << 1);  // set_tt, then set_fsp (ignored for Date)
  if (value == 0) { *bits = fsp_tt; return true; }     // Time::new(zero): every field 0 (Date also clears fsp)
  const uint64_t ymdhms = value >> 24, ymd = ymdhms >> 17, ym = ymd >> 5, hms = ymdhms & ((1u << 17) - 1);
  const uint64_t day = ymd & 31, month = ym % 13, year = ym / 13, second = hms & 63, minute = (hms >> 6) & 63, hour = hms >> 12, micro = value & ((1u << 24) - 1);
  *bits = ((year & 0x3fff) << 50) | ((month & 15) << 46) | ((day & 31) << 41) | ((hour & 31) << 36) | ((minute & 63) << 30) | ((second & 63) << 24) |
          ((micro & 0xfffff) << 4) | fsp_tt;
  return true;
}


// DateTime / Duration cells of an expression operand (decode_date_time_datum / decode_duration_datum, datum_codec.rs)
inline bool decode_time_datum(Slice d, const FieldType& ft, bool* is_null, int64_t* out, std::string* err) {
  *is_null = false; *out = 0;
  if (d.empty()) { *err = "Failed to decode datum flag"; return false; }
  Slice p = d.sub(1);
  uint64_t v;
  switch (d[0]) {
    case NIL_FLAG: *is_null = true; return true;
    case UINT_FLAG: if (p.n < 8) { *err = "unexpected eof"; return false; } v = get_u64_be(p.p); break;
    case VAR_UINT_FLAG: if (!decode_var_u64(p, &v)) { *err = "unexpected eof"; return false; } break;
    default: *err = "Unsupported datum flag " + std::to_string(d[0]) + " for DateTime vector."; return false;
  }
  uint64_t bits;
  if (!time_from_packed(v, ft.tp, ft.decimal, &bits, err)) return false;
  *out = (int64_t)bits;
  return true;
}
inline bool decode_duration_datum(Slice d, bool* is_null, int64_t* out, std::string* err) {
  *is_null = false; *out = 0;
  if (d.empty()) { *err = "Failed to decode datum flag"; return false; }
  Slice p = d.sub(1);
  switch (d[0]) {
    case NIL_FLAG: *is_null = true; return true;
    case DURATION_FLAG: if (p.n < 8) { *err = "unexpected eof"; return false; } *out = decode_i64(p.p); return true;
    case VAR_INT_FLAG: if (!decode_var_i64(p, out)) { *err = "unexpected eof"; return false; } return true;
    default: *err = "Unsupported datum flag " + std::to_string(d[0]) + " for Duration vector"; return false;
  }
}
inline bool et_int_like(EvalType et) { return et == ET_INT || et == ET_TIME || et == ET_DURATION; }

// lazy_column.rs:165-221 ensure_decoded: decode only logical rows, others become NULL
inline bool ensure_decoded(LazyColumn& c, const FieldType& ft, const std::vector<size_t>& logical_rows, std::string* err) {
  if (c.decoded) return true;
  EvalType et = eval_type_of(ft.tp);
  if (ft.tp == B2_TP_DATE || ft.tp == B2_TP_DATETIME) et = ET_TIME;
  if (ft.tp == B2_TP_DURATION) et = ET_DURATION;
  if (!et_int_like(et) && et != ET_REAL) { *err = "oracle decodes Int / Real / DateTime / Duration operands only"; return false; }
  size_t n = c.raw_offsets.size();
  std::vector<int64_t> iv; std::vector<double> fv; std::vector<uint8_t> nn(n, 0);
  if (et_int_like(et)) iv.assign(n, 0); else fv.assign(n, 0);
  for (size_t r : logical_rows) {
    bool is_null;
    if (et == ET_INT) { int64_t v; if (!decode_int_datum(c.raw_get(r), &is_null, &v, err)) return false; iv[r] = v; }
    else if (et == ET_TIME) { int64_t v; if (!decode_time_datum(c.raw_get(r), ft, &is_null, &v, err)) return false; iv[r] = v; }
    else if (et == ET_DURATION) { int64_t v; if (!decode_duration_datum(c.raw_get(r), &is_null, &v, err)) return false; iv[r] = v; }
    else { double v; if (!decode_real_datum(c.raw_get(r), ft.tp, &is_null, &v, err)) return false; fv[r] = v; }
    nn[r] = !is_null;
  }
  c.decoded = true; c.et = et; c.i64.swap(iv); c.f64.swap(fv); c.nn.swap(nn);
  c.raw_data.clear(); c.raw_offsets.clear();
  return true;
}

struct Batch {
  std::vector<LazyColumn> cols;
  std::vector<size_t> logical_rows;
  bool is_drained = false;
  Error err;  // error after the rows in this batch (interface.rs:229-235)
};

struct Executor {
  virtual ~Executor() {}
  virtual const std::vector<FieldType>& schema() const = 0;
  virtual void next_batch(size_t scan_rows, Batch* out) = 0;
  virtual ForwardScanner* scanner() { return nullptr; }
};

// ---- RangesScanner + TikvStorage over the forward scanner ----
struct RangesScanner {
  std::vector<std::pair<Bytes, Bytes>> ranges;  // raw [start, end)
  size_t cur = 0;
  bool in_range = false;
  ScannerConfig base_cfg;
  const CfView *w = nullptr, *l = nullptr, *d = nullptr;
  ForwardScanner fs;
  BackwardScanner bs;  // scan_backward_in_range (scanner.rs:25, 145): TableScan.desc
  bool desc = false;
  Statistics total;
  int met_newer = NEWER_UNKNOWN;
  bool met_lock = false;
  uint64_t rows = 0;

  void accumulate() {
    auto add = [](CfStatistics& a, const CfStatistics& b) {
      a.processed_keys += b.processed_keys; a.next += b.next; a.prev += b.prev; a.seek += b.seek; a.seek_for_prev += b.seek_for_prev; a.over_seek_bound += b.over_seek_bound;
    };
    Statistics& st = desc ? bs.statistics : fs.statistics;
    const int newer = desc ? bs.met_newer_ts_data : fs.met_newer_ts_data;
    add(total.write, st.write); add(total.lock, st.lock); add(total.data, st.data);
    total.processed_size += st.processed_size;
    if (newer == NEWER_MET) met_newer = NEWER_MET;
    else if (newer == NEWER_NOT_MET && met_newer == NEWER_UNKNOWN) met_newer = NEWER_NOT_MET;
    st = Statistics();
  }
  // returns 1 row, 0 drained, -1 error. key out = raw key (storage_impl.rs:82 Key::into_raw)
  int next(Bytes* raw_key, ScanOutput* so, Error* err) {
    for (;;) {
      if (!in_range) {
        if (cur >= ranges.size()) return 0;
        ScannerConfig cfg = base_cfg;
        cfg.has_lower = cfg.has_upper = true;
        cfg.lower_bound = key_from_raw(Slice(ranges[cur].first.data(), ranges[cur].first.size()));
        cfg.upper_bound = key_from_raw(Slice(ranges[cur].second.data(), ranges[cur].second.size()));
        if (desc) bs.init(cfg, w, l, d); else fs.init(cfg, w, l, d);
        in_range = true;
      }
      int r = desc ? bs.read_next(so, err) : fs.read_next(so, err);
      if (r < 0) { accumulate(); return -1; }
      if (r == 0) { accumulate(); in_range = false; cur++; continue; }
      if (decode_bytes(Slice(so->user_key.data(), so->user_key.size()), raw_key) == (size_t)-1) {
        *err = Error::make(B2_ERR_STORAGE, "invalid memcomparable user key");
        accumulate();
        return -1;
      }
      rows++;
      return 1;
    }
  }
};

// ---- table scan ----
struct TableScanExecutor : Executor {
  std::vector<FieldType> schema_;
  std::vector<Bytes> default_val;
  std::unordered_map<int64_t, size_t> column_id_index;
  std::vector<size_t> handle_indices;
  std::vector<uint8_t> is_column_filled;
  Bytes datum_buf_;
  RangesScanner rs;
  bool ended = false;
  // BatchIndexScanExecutor (index_scan_executor.rs:47-170), old-format values only (value.len() <= 9, :375-379): the
  // index columns come from the key's datums, the int handle from the key tail (non-unique index) or from the value
  // (unique index); the PK handle, if requested, is the last column before the optional physical-table-id column.
  bool is_index = false;
  size_t idx_cols_without_handle = 0;
  bool idx_decode_int_handle = false, idx_physical_table_id = false;

  void init_index(const b2_executor_desc& d) {
    is_index = true;
    for (uint32_t i = 0; i < d.n_columns; ++i) { FieldType ft; ft.tp = d.columns[i].tp; ft.flag = d.columns[i].flag; ft.decimal = d.columns[i].decimal; schema_.push_back(ft); }
    size_t n = d.n_columns;
    idx_physical_table_id = n > 0 && d.columns[n - 1].col_id == B2_EXTRA_PHYSICAL_TABLE_ID_COL_ID;
    size_t tail = idx_physical_table_id ? 1 : 0;
    idx_decode_int_handle = n > tail && d.columns[n - 1 - tail].pk_handle;
    idx_cols_without_handle = n - tail - (idx_decode_int_handle ? 1 : 0);
    is_column_filled.assign(n, 0);
  }

  // process_kv_pair :363-380 -> process_old_collation_kv :514-574
  bool process_index_kv(Slice key, Slice value, std::vector<LazyColumn>& columns, std::string* err) {
    if (key.n < 1 || key[0] != 't') { *err = "record or index key expected"; return false; }  // check_index_key table.rs:114-140
    if (key.n < 11) { *err = "unexpected eof"; return false; }
    if (key[9] != '_' || key[10] != 'i') { *err = "expected key sep type _i"; return false; }
    if (key.n < 19) { *err = "unexpected eof"; return false; }
    if (value.n > 9) { *err = "index value in the new (restored-data) layout is not restated"; return false; }
    Slice payload = key.sub(19);
    for (size_t i = 0; i < idx_cols_without_handle; ++i) {  // extract_columns_from_datum_format :493-506
      if (payload.empty()) { *err = std::to_string(i) + "th column is missing value"; return false; }
      size_t dl = split_datum(payload, err);
      if (!dl) return false;
      columns[i].raw_push(payload.sub(0, dl));
      payload = payload.sub(dl);
    }
    if (idx_decode_int_handle) {
      int64_t handle;
      if (payload.empty()) {  // unique index: decode_int_handle_from_value :406-412 (plain big-endian u64)
        if (value.n < 8) { *err = "Failed to decode handle in value as i64"; return false; }
        uint64_t u = 0; for (int b = 0; b < 8; ++b) u = (u << 8) | value[b];
        handle = (int64_t)u;
      } else {  // non-unique index: decode_int_handle_from_key :451-471
        uint8_t flag = payload[0];
        if (flag != INT_FLAG && flag != UINT_FLAG) { *err = "Unexpected handle flag " + std::to_string(flag); return false; }
        if (payload.n < 9) { *err = flag == INT_FLAG ? "Failed to decode handle in key as i64" : "Failed to decode handle in key as u64"; return false; }
        uint64_t u = 0; for (int b = 0; b < 8; ++b) u = (u << 8) | payload[1 + b];
        handle = flag == INT_FLAG ? (int64_t)(u ^ 0x8000000000000000ull) : (int64_t)u;
      }
      columns[idx_cols_without_handle].push_int(true, handle);
    }
    if (idx_physical_table_id) {  // process_physical_table_id_column: the table id of the key prefix
      int64_t tid; const char* e = decode_table_id(key, &tid);
      if (e) { *err = e; return false; }
      columns[columns.size() - 1].push_int(true, tid);
    }
    return true;
  }

  void init(const b2_executor_desc& d) {
    if (d.tp == B2_EXEC_INDEX_SCAN) { init_index(d); return; }
    for (uint32_t i = 0; i < d.n_columns; ++i) {
      const b2_column_info& ci = d.columns[i];
      FieldType ft; ft.tp = ci.tp; ft.flag = ci.flag; ft.decimal = ci.decimal;
      schema_.push_back(ft);
      default_val.push_back(ci.default_val ? Bytes(ci.default_val, ci.default_val + ci.default_len) : Bytes());
      if (ci.pk_handle) handle_indices.push_back(i);
      else column_id_index[ci.col_id] = i;  // last one wins (table_scan_executor.rs:90-94)
    }
    is_column_filled.assign(d.n_columns, 0);
    rs.base_cfg.load_commit_ts = column_id_index.count(B2_EXTRA_COMMIT_TS_COL_ID) > 0;
  }
  const std::vector<FieldType>& schema() const override { return schema_; }
  ForwardScanner* scanner() override { return &rs.fs; }

  bool is_decoded_col(size_t i) const {
    if (is_index) return i >= idx_cols_without_handle;
    for (size_t h : handle_indices) if (h == i) return true;
    auto a = column_id_index.find(B2_EXTRA_PHYSICAL_TABLE_ID_COL_ID);
    if (a != column_id_index.end() && a->second == i) return true;
    auto b = column_id_index.find(B2_EXTRA_COMMIT_TS_COL_ID);
    if (b != column_id_index.end() && b->second == i) return true;
    return false;
  }

  bool process_v1(Slice value, std::vector<LazyColumn>& columns, size_t* decoded_columns, std::string* err) {  // :200-247
    size_t columns_len = columns.size();
    Slice remaining = value;
    while (!remaining.empty() && *decoded_columns < columns_len) {
      if (remaining[0] != VAR_INT_FLAG) { *err = "Unable to decode row: column id must be VAR_INT"; return false; }
      remaining = remaining.sub(1);
      int64_t column_id;
      size_t n = decode_var_i64(remaining, &column_id);
      if (!n) { *err = "unexpected eof"; return false; }
      remaining = remaining.sub(n);
      size_t dl = split_datum(remaining, err);
      if (!dl) return false;
      Slice val = remaining.sub(0, dl);
      auto it = column_id_index.find(column_id);
      if (it != column_id_index.end()) {
        size_t index = it->second;
        if (!is_column_filled[index]) {
          columns[index].raw_push(val);
          (*decoded_columns)++;
          is_column_filled[index] = 1;
        }
      }
      remaining = remaining.sub(dl);
    }
    return true;
  }

  // RowSlice::from_bytes + search (row_slice.rs:74-166, 330-357) + write_v2_as_datum (compat_v1.rs:13-129)
  bool process_v2(Slice value, std::vector<LazyColumn>& columns, size_t* decoded_columns, std::string* err) {  // :249-281
    if (value.n < 6) { *err = "unexpected eof"; return false; }
    uint8_t flags = value[1];
    bool is_big = flags & 1, with_checksum = flags & 2;
    size_t non_null_cnt = value[2] | (value[3] << 8), null_cnt = value[4] | (value[5] << 8);
    size_t id_w = is_big ? 4 : 1, off_w = is_big ? 4 : 2;
    Slice data = value.sub(6);
    if (data.n < non_null_cnt * id_w) { *err = "unexpected eof"; return false; }
    Slice non_null_ids = data.sub(0, non_null_cnt * id_w); data = data.sub(non_null_cnt * id_w);
    if (data.n < null_cnt * id_w) { *err = "unexpected eof"; return false; }
    Slice null_ids = data.sub(0, null_cnt * id_w); data = data.sub(null_cnt * id_w);
    if (data.n < non_null_cnt * off_w) { *err = "unexpected eof"; return false; }
    Slice offsets = data.sub(0, non_null_cnt * off_w); data = data.sub(non_null_cnt * off_w);
    Slice values = data;
    auto get_id = [&](Slice ids, size_t i) -> uint32_t { return is_big ? (uint32_t)(ids[4 * i] | (ids[4 * i + 1] << 8) | (ids[4 * i + 2] << 16) | ((uint32_t)ids[4 * i + 3] << 24)) : ids[i]; };
    auto get_off = [&](size_t i) -> size_t { return is_big ? (size_t)(offsets[4 * i] | (offsets[4 * i + 1] << 8) | (offsets[4 * i + 2] << 16) | ((uint32_t)offsets[4 * i + 3] << 24)) : (size_t)(offsets[2 * i] | (offsets[2 * i + 1] << 8)); };
    if (with_checksum) {
      size_t last = non_null_cnt == 0 ? 0 : get_off(non_null_cnt - 1);
      if (last > values.n) { *err = "row v2 checksum cut out of range (panic)"; return false; }
      size_t ck = values.n - last;
      if (ck != 5 && ck != 9) { *err = "row v2 checksum bytes must be 5 or 9 (assert)"; return false; }
      values = values.sub(0, last);
    }
    // LeBytes::binary_search (row_slice.rs:330-357)
    auto bsearch = [&](Slice ids, size_t cnt, uint32_t v, size_t* idx) -> bool {
      if (cnt == 0) return false;
      size_t size = cnt, base = 0, steps = 20;
      while (steps > 0 && size > 1) { size_t half = size / 2, mid = base + half; if (!(get_id(ids, mid) > v)) base = mid; size -= half; steps--; }
      if (get_id(ids, base) == v) { *idx = base; return true; }
      return false;
    };
    int64_t upper = is_big ? 0xffffffffll : 0xffll;
    for (auto& kv : column_id_index) {
      int64_t col_id = kv.first; size_t idx = kv.second;
      if (is_column_filled[idx]) continue;
      bool id_valid = col_id > 0 && col_id <= upper;
      size_t pos;
      if (id_valid && bsearch(non_null_ids, non_null_cnt, is_big ? (uint32_t)col_id : (uint8_t)col_id, &pos)) {
        size_t end = get_off(pos), start = pos > 0 ? get_off(pos - 1) : 0;
        if (start > end || end > values.n) { *err = "row v2 value slice out of range (panic)"; return false; }
        Slice src = values.sub(start, end - start);
        Bytes& datum = datum_buf_;  // scratch reused across cells
        datum.clear();
        const FieldType& ft = schema_[idx];
        switch (ft.tp) {
          case B2_TP_TINY: case B2_TP_SHORT: case B2_TP_INT24: case B2_TP_LONG: case B2_TP_LONGLONG: case B2_TP_YEAR:
          case B2_TP_DATE: case B2_TP_DATETIME: case B2_TP_TIMESTAMP: case B2_TP_ENUM: case B2_TP_BIT: case B2_TP_SET: {
            bool as_unsigned = (ft.tp == B2_TP_YEAR) ? false : ((ft.tp == B2_TP_TINY || ft.tp == B2_TP_SHORT || ft.tp == B2_TP_INT24 || ft.tp == B2_TP_LONG || ft.tp == B2_TP_LONGLONG) ? ft.is_unsigned() : true);
            uint64_t u;
            switch (src.n) {
              case 1: u = as_unsigned ? (uint64_t)src[0] : (uint64_t)(int64_t)(int8_t)src[0]; break;
              case 2: { uint16_t x = (uint16_t)(src[0] | (src[1] << 8)); u = as_unsigned ? (uint64_t)x : (uint64_t)(int64_t)(int16_t)x; break; }
              case 4: { uint32_t x = (uint32_t)(src[0] | (src[1] << 8) | (src[2] << 16) | ((uint32_t)src[3] << 24)); u = as_unsigned ? (uint64_t)x : (uint64_t)(int64_t)(int32_t)x; break; }
              case 8: u = get_u64_le(src.p); break;
              default: *err = as_unsigned ? "Failed to decode row v2 data as u64" : "Failed to decode row v2 data as i64"; return false;
            }
            if (as_unsigned) { datum.push_back(UINT_FLAG); put_u64_be(datum, u); }
            else { datum.push_back(INT_FLAG); encode_i64(datum, (int64_t)u); }
            break;
          }
          case B2_TP_FLOAT: case B2_TP_DOUBLE: datum.push_back(FLOAT_FLAG); datum.insert(datum.end(), src.p, src.p + src.n); break;
          case B2_TP_NEWDECIMAL: datum.push_back(DECIMAL_FLAG); datum.insert(datum.end(), src.p, src.p + src.n); break;
          case B2_TP_JSON: datum.push_back(JSON_FLAG); datum.insert(datum.end(), src.p, src.p + src.n); break;
          case B2_TP_DURATION: {
            if (src.n != 1 && src.n != 2 && src.n != 4 && src.n != 8) { *err = "Failed to decode row v2 data as i64"; return false; }
            int64_t v = src.n == 1 ? (int8_t)src[0] : src.n == 2 ? (int16_t)(src[0] | (src[1] << 8)) : src.n == 4 ? (int32_t)(src[0] | (src[1] << 8) | (src[2] << 16) | ((uint32_t)src[3] << 24)) : (int64_t)get_u64_le(src.p);
            datum.push_back(DURATION_FLAG); encode_i64(datum, v);
            break;
          }
          case B2_TP_VARCHAR: case B2_TP_VARSTRING: case B2_TP_STRING: case B2_TP_BLOB: case 0xf9: case 0xfa: case 0xfb: case 0xff:
            datum.push_back(COMPACT_BYTES_FLAG); encode_var_i64(datum, (int64_t)src.n); datum.insert(datum.end(), src.p, src.p + src.n); break;
          case B2_TP_NULL: datum.push_back(NIL_FLAG); break;
          default: *err = "Unsupported FieldType"; return false;
        }
        columns[idx].raw_push(Slice(datum.data(), datum.size()));
        (*decoded_columns)++;
        is_column_filled[idx] = 1;
      } else if (id_valid && bsearch(null_ids, null_cnt, is_big ? (uint32_t)col_id : (uint8_t)col_id, &pos)) {
        uint8_t nil = NIL_FLAG;
        columns[idx].raw_push(Slice(&nil, 1));
        (*decoded_columns)++;
        is_column_filled[idx] = 1;
      }
    }
    return true;
  }

  bool process_kv_pair(Slice key, Slice value, std::vector<LazyColumn>& columns, bool has_commit_ts, uint64_t commit_ts, std::string* err) {  // :365-475
    size_t columns_len = schema_.size();
    size_t decoded_columns = 0;
    if (value.empty() || (value.n == 1 && value[0] == NIL_FLAG)) {
    } else if (value[0] == 128) { if (!process_v2(value, columns, &decoded_columns, err)) return false; }
    else { if (!process_v1(value, columns, &decoded_columns, err)) return false; }
    if (!handle_indices.empty()) {
      int64_t handle;
      const char* e = decode_int_handle(key, &handle);
      if (e) { *err = e; return false; }
      for (size_t hi : handle_indices) if (!is_column_filled[hi]) { columns[hi].push_int(true, handle); decoded_columns++; is_column_filled[hi] = 1; }
    } else {
      const char* e = check_record_key(key);
      if (e) { *err = e; return false; }
    }
    auto pt = column_id_index.find(B2_EXTRA_PHYSICAL_TABLE_ID_COL_ID);
    if (pt != column_id_index.end()) {
      int64_t tid; const char* e = decode_table_id(key, &tid);
      if (e) { *err = e; return false; }
      columns[pt->second].push_int(true, tid);
      is_column_filled[pt->second] = 1;
    }
    auto ct = column_id_index.find(B2_EXTRA_COMMIT_TS_COL_ID);
    if (ct != column_id_index.end()) {
      if (!has_commit_ts) { *err = "Query asks for _tidb_commit_ts, but the data is missing"; return false; }
      columns[ct->second].push_int(true, (int64_t)commit_ts);
      is_column_filled[ct->second] = 1;
    }
    for (size_t i = 0; i < columns_len; ++i) {
      if (!is_column_filled[i]) {
        if (!default_val[i].empty()) columns[i].raw_push(Slice(default_val[i].data(), default_val[i].size()));
        else if (!(schema_[i].flag & B2_FLAG_NOT_NULL)) { uint8_t nil = NIL_FLAG; columns[i].raw_push(Slice(&nil, 1)); }
        else { *err = "Data is corrupted, missing data for NOT NULL column (offset = " + std::to_string(i) + ")"; return false; }
      } else is_column_filled[i] = 0;
    }
    return true;
  }

  void next_batch(size_t scan_rows, Batch* out) override {  // scan_executor.rs:226-261 + fill_column_vec :114-169
    out->cols.assign(schema_.size(), LazyColumn());
    for (size_t i = 0; i < schema_.size(); ++i) if (is_decoded_col(i)) { out->cols[i].decoded = true; out->cols[i].et = ET_INT; }
    out->is_drained = false; out->err = Error();
    Bytes raw_key; ScanOutput so;  // buffers reused across rows
    for (size_t i = 0; i < scan_rows; ++i) {
      Error e;
      int r = rs.next(&raw_key, &so, &e);
      if (r < 0) { out->err = e; break; }
      if (r == 0) { out->is_drained = true; break; }
      std::string perr;
      const Slice rk(raw_key.data(), raw_key.size()), rv(so.value.data(), so.value.size());
      if (!(is_index ? process_index_kv(rk, rv, out->cols, &perr) : process_kv_pair(rk, rv, out->cols, so.has_commit_ts, so.commit_ts, &perr))) {
        // truncate_into_equal_length (lazy_column_vec.rs:210-220)
        size_t m = (size_t)-1;
        for (auto& c : out->cols) m = std::min(m, c.len());
        for (auto& c : out->cols) { if (c.decoded) { c.i64.resize(m); c.nn.resize(m); } else c.raw_truncate(m); }
        std::fill(is_column_filled.begin(), is_column_filled.end(), 0);
        out->err = Error::make(B2_ERR_CORRUPTED, perr);
        break;
      }
    }
    size_t n = out->cols.empty() ? 0 : out->cols[0].len();
    out->logical_rows.resize(n);
    for (size_t i = 0; i < n; ++i) out->logical_rows[i] = i;
  }
};

// ---- RPN evaluation ----
struct Val {  // one stack node restricted to the batch's logical rows (index j <-> logical_rows[j])
  bool scalar = false;
  EvalType et = ET_INT;
  bool is_unsigned = false;
  bool s_null = true; int64_t s_i = 0; double s_f = 0;
  std::vector<int64_t> i; std::vector<double> f; std::vector<uint8_t> nn;
  std::vector<Bytes> b; Bytes s_b;  // ET_BYTES
  const Bytes& bytes_at(size_t j) const { return scalar ? s_b : b[j]; }
  std::vector<Decimal> d; Decimal s_d;  // ET_DEC
  const Decimal& dec_at(size_t j) const { return scalar ? s_d : d[j]; }
  bool null_at(size_t j) const { return scalar ? s_null : !nn[j]; }
  int64_t int_at(size_t j) const { return scalar ? s_i : i[j]; }
  double real_at(size_t j) const { return scalar ? s_f : f[j]; }
};

inline Error overflow_err(const char* tp) { return Error::make(B2_ERR_EVALUATE, std::string(tp) + " value is out of range", B2_MYSQL_ERR_DATA_OUT_OF_RANGE); }

inline int cmp_int(int64_t a, bool au, int64_t b, bool bu) {  // impl_compare.rs:63-149
  if (!au && !bu) return a < b ? -1 : a > b;
  if (au && bu) return (uint64_t)a < (uint64_t)b ? -1 : (uint64_t)a > (uint64_t)b;
  if (au && !bu) { if (b < 0 || (uint64_t)a > (uint64_t)INT64_MAX) return 1; return a < b ? -1 : a > b; }
  if (a < 0 || (uint64_t)b > (uint64_t)INT64_MAX) return -1;
  return a < b ? -1 : a > b;
}

struct ExprCtx { const std::vector<FieldType>* schema; Batch* batch; };

inline bool dec_read(Slice s, Decimal* out, std::string* err);  // orc_chunk.h (decimal.rs:2204-2289)
inline bool decode_decimal_datum(Slice d, bool* is_null, Decimal* out, std::string* err) {  // decode_decimal_datum, datum_codec.rs
  *is_null = false;
  if (d.empty()) { *err = "Failed to decode datum flag"; return false; }
  if (d[0] == NIL_FLAG) { *is_null = true; *out = dec_zero(); return true; }
  if (d[0] != DECIMAL_FLAG) { *err = "Unsupported datum flag " + std::to_string(d[0]) + " for Decimal vector"; return false; }
  return dec_read(d.sub(1), out, err);
}

inline bool tp_is_bytes(int tp) {  // EvalType::Bytes (def/eval_type.rs:53-95)
  switch (tp) { case B2_TP_VARCHAR: case B2_TP_VARSTRING: case B2_TP_STRING: case B2_TP_BLOB: case 0xf9: case 0xfa: case 0xfb: case 0xff: return true; default: return false; }
}
// decode_bytes_datum (datum_codec.rs): NIL, compact bytes, memcomparable bytes
inline bool decode_bytes_datum(Slice d, bool* is_null, Bytes* out, std::string* err) {
  *is_null = false; out->clear();
  if (d.empty()) { *err = "Failed to decode datum flag"; return false; }
  Slice p = d.sub(1);
  switch (d[0]) {
    case NIL_FLAG: *is_null = true; return true;
    case COMPACT_BYTES_FLAG: {
      int64_t vn; size_t n = decode_var_i64(p, &vn);
      if (!n || vn < 0 || p.n - n < (size_t)vn) { *err = "unexpected eof"; return false; }
      out->assign(p.p + n, p.p + n + vn);
      return true;
    }
    case BYTES_FLAG: if (decode_bytes(p, out) == (size_t)-1) { *err = "unexpected eof"; return false; } return true;
    default: *err = "Unsupported datum flag " + std::to_string(d[0]) + " for Bytes vector"; return false;
  }
}

// ---- LIKE (impl_like.rs:7-74) ----
// Charset::decode_one: binary = one byte (charset.rs:17-24); utf8mb4 = core::str::next_code_point (charset.rs:43-54: the
// lead byte decides the length, nothing is validated)
inline size_t like_decode_one(const uint8_t* s, size_t n, bool utf8, uint32_t* code) {
  if (n == 0) return 0;
  uint32_t x = s[0];
  if (!utf8 || x < 128) { *code = x; return 1; }
  uint32_t y = n > 1 ? (s[1] & 0x3f) : 0, ch = ((x & 0x1f) << 6) | y;
  size_t len = 2;
  if (x >= 0xe0) {
    uint32_t yz = (y << 6) | (n > 2 ? (s[2] & 0x3f) : 0);
    ch = ((x & 0x1f) << 12) | yz; len = 3;
    if (x >= 0xf0) { ch = ((x & 7) << 18) | (yz << 6) | (n > 3 ? (s[3] & 0x3f) : 0); len = 4; }
  }
  *code = ch;
  return std::min(len, n);
}
// like::<C, CS> for a collator C whose sort_compare(.., force_no_pad = true) of two single characters is byte equality
// (CollatorBinary, CollatorUtf8Mb4Bin, CollatorUtf8Mb4BinNoPadding)
inline bool like_eval(const Bytes& target, const Bytes& pattern, int64_t escape_i, bool utf8) {
  const uint32_t escape = (uint32_t)escape_i;
  size_t px = 0, tx = 0, next_px = 0, next_tx = 0;
  while (px < pattern.size() || tx < target.size()) {
    uint32_t code = 0, tc = 0;
    size_t poff = like_decode_one(pattern.data() + px, pattern.size() - px, utf8, &code);
    if (poff) {
      if (code == '_') {
        size_t toff = like_decode_one(target.data() + tx, target.size() - tx, utf8, &tc);
        if (toff) { px += poff; tx += toff; continue; }
      } else if (code == '%') {
        px += poff; next_px = px;
        if (next_px >= pattern.size()) return true;
        next_tx = tx;
        continue;
      } else {
        bool brk = false;
        if (code == escape && px + poff < pattern.size()) {
          px += poff;
          uint32_t c2;
          poff = like_decode_one(pattern.data() + px, pattern.size() - px, utf8, &c2);
          if (!poff) brk = true;
        }
        if (brk) break;
        size_t toff = like_decode_one(target.data() + tx, target.size() - tx, utf8, &tc);
        if (toff && toff == poff && memcmp(target.data() + tx, pattern.data() + px, toff) == 0) { tx += toff; px += poff; continue; }
      }
    }
    if (0 < next_px && next_tx < target.size()) {
      size_t toff = like_decode_one(target.data() + next_tx, target.size() - next_tx, utf8, &tc);
      next_tx += toff ? toff : 1;
      px = next_px; tx = next_tx;
      continue;
    }
    return false;
  }
  return true;
}
// Collation::from_i32 (field_type.rs:130-146) restricted to what like_eval covers: charset 0 binary, 1 utf8mb4; false = other
inline bool like_collation(int n, int* charset) {
  switch (n) {
    case -63: case 63: case 47: *charset = 0; return true;
    case -46: case -83: case -65: case -309: *charset = 1; return true;
    default: if (n >= 0) { *charset = 1; return true; } return false;
  }
}
// EvalContext::warnings (expr/ctx.rs:180-215) of the request running on this thread: only "Division by 0" can occur here
inline uint64_t& warning_count() { static thread_local uint64_t n = 0; return n; }

inline bool rpn_eval(const b2_rpn_expr& e, ExprCtx& cx, Val* result, Error* err) {
  size_t n = cx.batch->logical_rows.size();
  std::vector<Val> st;
  for (uint32_t k = 0; k < e.n_nodes; ++k) {
    const b2_rpn_node& nd = e.nodes[k];
    if (nd.kind == B2_RPN_CONST_DECIMAL) {  // ExprType::MysqlDecimal: prec, frac, binary decimal
      Val v; v.scalar = true; v.s_null = false; v.et = ET_DEC;
      std::string perr;
      if (!dec_read(Slice((const uint8_t*)(uintptr_t)nd.i64, (size_t)nd.n_args), &v.s_d, &perr)) { *err = Error::make(B2_ERR_INVALID_ARG, perr); return false; }
      st.push_back(std::move(v));
    } else if (nd.kind == B2_RPN_CONST_BYTES) {
      Val v; v.scalar = true; v.s_null = false; v.et = ET_BYTES;
      const uint8_t* src = (const uint8_t*)(uintptr_t)nd.i64;
      if (nd.n_args > 0) v.s_b.assign(src, src + nd.n_args);
      st.push_back(std::move(v));
    } else if (nd.kind == B2_RPN_CONST_TIME || nd.kind == B2_RPN_CONST_DURATION) {  // ExprType::MysqlTime / MysqlDuration constants
      Val v; v.scalar = true; v.s_null = false;
      if (nd.kind == B2_RPN_CONST_TIME) {
        uint64_t bits; std::string perr;
        if (!time_from_packed((uint64_t)nd.i64, nd.field_tp, 0, &bits, &perr)) { *err = Error::make(B2_ERR_UNSUPPORTED, perr); return false; }
        v.et = ET_TIME; v.s_i = (int64_t)bits;
      } else { v.et = ET_DURATION; v.s_i = nd.i64; }
      st.push_back(std::move(v));
    } else if (nd.kind == B2_RPN_CONST_NULL || nd.kind == B2_RPN_CONST_INT || nd.kind == B2_RPN_CONST_UINT || nd.kind == B2_RPN_CONST_REAL) {
      Val v; v.scalar = true;
      v.et = nd.kind == B2_RPN_CONST_REAL ? ET_REAL : (nd.kind == B2_RPN_CONST_NULL ? eval_type_of(nd.field_tp) : ET_INT);
      if (nd.kind == B2_RPN_CONST_NULL && tp_is_bytes(nd.field_tp)) v.et = ET_BYTES;
      else if (nd.kind == B2_RPN_CONST_NULL && nd.field_tp == B2_TP_NEWDECIMAL) { v.et = ET_DEC; v.s_d = dec_zero(); }
      else if (nd.kind == B2_RPN_CONST_NULL && (nd.field_tp == B2_TP_DATE || nd.field_tp == B2_TP_DATETIME)) v.et = ET_TIME;
      else if (nd.kind == B2_RPN_CONST_NULL && nd.field_tp == B2_TP_DURATION) v.et = ET_DURATION;
      else if (v.et != ET_REAL) v.et = ET_INT;
      v.is_unsigned = (nd.field_flag & B2_FLAG_UNSIGNED) || nd.kind == B2_RPN_CONST_UINT;
      v.s_null = nd.kind == B2_RPN_CONST_NULL; v.s_i = nd.i64; v.s_f = nd.f64;
      st.push_back(std::move(v));
    } else if (nd.kind == B2_RPN_COLUMN_REF) {
      size_t ci = (size_t)nd.i64;
      if (ci >= cx.batch->cols.size()) { *err = Error::make(B2_ERR_INVALID_ARG, "column offset out of range"); return false; }
      std::string perr;
      if ((*cx.schema)[ci].tp == B2_TP_NEWDECIMAL && !cx.batch->cols[ci].decoded) {  // Decimal operand (the column stays Raw, as for bytes below)
        const LazyColumn& c = cx.batch->cols[ci];
        Val v; v.et = ET_DEC; v.nn.resize(n); v.d.resize(n);
        for (size_t j = 0; j < n; ++j) {
          bool is_null;
          if (!decode_decimal_datum(c.raw_get(cx.batch->logical_rows[j]), &is_null, &v.d[j], &perr)) { *err = Error::make(B2_ERR_CORRUPTED, perr); return false; }
          v.nn[j] = !is_null;
        }
        st.push_back(std::move(v));
        continue;
      }
      if (tp_is_bytes((*cx.schema)[ci].tp) && !cx.batch->cols[ci].decoded) {
        // Bytes operand.  (The reference decodes the whole column here, lazy_column.rs:165-221; the cells it then sends are the
        // same bytes, so this restatement reads the operands from the raw datums and leaves the column Raw.)
        const LazyColumn& c = cx.batch->cols[ci];
        Val v; v.et = ET_BYTES; v.nn.resize(n); v.b.resize(n);
        for (size_t j = 0; j < n; ++j) {
          bool is_null;
          if (!decode_bytes_datum(c.raw_get(cx.batch->logical_rows[j]), &is_null, &v.b[j], &perr)) { *err = Error::make(B2_ERR_CORRUPTED, perr); return false; }
          v.nn[j] = !is_null;
        }
        st.push_back(std::move(v));
        continue;
      }
      if (!ensure_decoded(cx.batch->cols[ci], (*cx.schema)[ci], cx.batch->logical_rows, &perr)) { *err = Error::make(B2_ERR_CORRUPTED, perr); return false; }
      const LazyColumn& c = cx.batch->cols[ci];
      Val v; v.et = c.et; v.is_unsigned = (*cx.schema)[ci].is_unsigned();
      v.nn.resize(n);
      if (et_int_like(c.et)) v.i.resize(n); else v.f.resize(n);
      for (size_t j = 0; j < n; ++j) { size_t r = cx.batch->logical_rows[j]; v.nn[j] = c.nn[r]; if (et_int_like(c.et)) v.i[j] = c.i64[r]; else v.f[j] = c.f64[r]; }
      st.push_back(std::move(v));
    } else if (nd.kind == B2_RPN_FN) {
      int na = nd.n_args;
      if (nd.sig == B2_SIG_LIKE) {  // like::<C, CS>(target, pattern, escape), lib.rs:99-135 map_like_sig picks C and CS
        if ((int)st.size() < 3 || na != 3) { *err = Error::make(B2_ERR_INVALID_ARG, "bad rpn arity"); return false; }
        std::vector<Val> args(st.end() - 3, st.end());
        st.resize(st.size() - 3);
        if (args[0].et != ET_BYTES || args[1].et != ET_BYTES || args[2].et != ET_INT) { *err = Error::make(B2_ERR_INVALID_ARG, "argument eval type does not match the function"); return false; }
        // the operands' own collations: the last nodes of the target and pattern subtrees
        int kk = (int)k - 1;
        auto skip = [&](int& q) { int want = 1; while (want > 0 && q >= 0) { const b2_rpn_node& z = e.nodes[q]; want += (z.kind == B2_RPN_FN ? z.n_args : 0) - 1; --q; } };
        skip(kk); const int pat = kk; skip(kk); const int tgt = kk;
        int cs_r, cs_t, cs_p;
        if (tgt < 0 || !like_collation(nd.collation, &cs_r) || !like_collation(e.nodes[tgt].collation, &cs_t) || !like_collation(e.nodes[pat].collation, &cs_p)) {
          *err = Error::make(B2_ERR_UNSUPPORTED, "oracle restates LIKE for the binary and *_bin collations only"); return false;
        }
        const bool utf8 = (cs_t == cs_p ? cs_t : cs_r) == 1;
        Val r; r.et = ET_INT; r.is_unsigned = false; r.nn.assign(n, 0); r.i.assign(n, 0);
        for (size_t j = 0; j < n; ++j) {
          if (args[0].null_at(j) || args[1].null_at(j) || args[2].null_at(j)) continue;
          r.nn[j] = 1; r.i[j] = like_eval(args[0].bytes_at(j), args[1].bytes_at(j), args[2].int_at(j), utf8);
        }
        st.push_back(std::move(r));
        continue;
      }
      {  // Decimal comparisons, IN, IS NULL (impl_compare.rs:63-240 over `Ord for Decimal`, decimal.rs:2323-2338)
        const int sig = nd.sig;
        const bool dcmp = sig >= 100 && sig < 170 && sig % 10 == 2, din = sig == B2_SIG_IN_DECIMAL, dnull = sig == B2_SIG_DECIMAL_IS_NULL;
        if (dcmp || din || dnull) {
          if ((int)st.size() < na || na < 1 || (dcmp && na != 2) || (dnull && na != 1)) { *err = Error::make(B2_ERR_INVALID_ARG, "bad rpn arity"); return false; }
          std::vector<Val> args(st.end() - na, st.end());
          st.resize(st.size() - na);
          for (auto& a : args) if (a.et != ET_DEC) { *err = Error::make(B2_ERR_INVALID_ARG, "argument eval type does not match the function"); return false; }
          Val r; r.et = ET_INT; r.is_unsigned = false; r.nn.assign(n, 0); r.i.assign(n, 0);
          for (size_t j = 0; j < n; ++j) {
            if (dnull) { r.nn[j] = 1; r.i[j] = args[0].null_at(j); continue; }
            if (din) {
              if (args[0].null_at(j)) continue;
              bool hit = false, default_null = false;
              for (int i = 1; i < na; ++i) { if (args[i].null_at(j)) { default_null = true; continue; } hit |= dec_cmp(args[0].dec_at(j), args[i].dec_at(j)) == 0; }
              if (hit) { r.nn[j] = 1; r.i[j] = 1; } else if (!default_null) { r.nn[j] = 1; r.i[j] = 0; }
              continue;
            }
            const bool an = args[0].null_at(j), bn = args[1].null_at(j), nulleq = sig / 10 * 10 == B2_SIG_NULLEQ_INT;
            if (an && bn) { if (nulleq) { r.nn[j] = 1; r.i[j] = 1; } continue; }
            if (an || bn) { if (nulleq) { r.nn[j] = 1; r.i[j] = 0; } continue; }
            const int c = dec_cmp(args[0].dec_at(j), args[1].dec_at(j));
            bool res;
            switch (sig / 10 * 10) {
              case B2_SIG_LT_INT: res = c < 0; break; case B2_SIG_LE_INT: res = c <= 0; break;
              case B2_SIG_GT_INT: res = c > 0; break; case B2_SIG_GE_INT: res = c >= 0; break;
              case B2_SIG_NE_INT: res = c != 0; break; default: res = c == 0; break;
            }
            r.nn[j] = 1; r.i[j] = res;
          }
          st.push_back(std::move(r));
          continue;
        }
      }
      {  // DateTime / Duration comparisons, IN, IS NULL (impl_compare.rs:63-240 over `Ord for Time`, mysql/time/mod.rs:2814-2840:
         // set_fsp_tt(0) on both sides, then the u64 bit fields compare; `Ord for Duration`: the nanoseconds compare)
        const int sig = nd.sig;
        const bool tcmp = sig >= 100 && sig < 170 && (sig % 10 == 4 || sig % 10 == 5);
        const bool tin = sig == B2_SIG_IN_TIME || sig == B2_SIG_IN_DURATION, tnull = sig == B2_SIG_TIME_IS_NULL || sig == B2_SIG_DURATION_IS_NULL;
        if (tcmp || tin || tnull) {
          const EvalType need = (tcmp ? sig % 10 == 4 : (sig == B2_SIG_IN_TIME || sig == B2_SIG_TIME_IS_NULL)) ? ET_TIME : ET_DURATION;
          if ((int)st.size() < na || na < 1 || (tcmp && na != 2) || (tnull && na != 1)) { *err = Error::make(B2_ERR_INVALID_ARG, "bad rpn arity"); return false; }
          std::vector<Val> args(st.end() - na, st.end());
          st.resize(st.size() - na);
          for (auto& a : args) if (a.et != need) { *err = Error::make(B2_ERR_INVALID_ARG, "argument eval type does not match the function"); return false; }
          auto key = [&](const Val& v, size_t j) -> int64_t { return need == ET_TIME ? (int64_t)((uint64_t)v.int_at(j) & ~15ull) : v.int_at(j); };
          auto cmp = [&](int64_t x, int64_t y) -> int { return need == ET_TIME ? ((uint64_t)x < (uint64_t)y ? -1 : (uint64_t)x > (uint64_t)y) : (x < y ? -1 : x > y); };
          Val r; r.et = ET_INT; r.is_unsigned = false; r.nn.assign(n, 0); r.i.assign(n, 0);
          for (size_t j = 0; j < n; ++j) {
            if (tnull) { r.nn[j] = 1; r.i[j] = args[0].null_at(j); continue; }
            if (tin) {
              if (args[0].null_at(j)) continue;
              bool hit = false, default_null = false;
              for (int i = 1; i < na; ++i) { if (args[i].null_at(j)) { default_null = true; continue; } hit |= key(args[0], j) == key(args[i], j); }
              if (hit) { r.nn[j] = 1; r.i[j] = 1; } else if (!default_null) { r.nn[j] = 1; r.i[j] = 0; }
              continue;
            }
            const bool an = args[0].null_at(j), bn = args[1].null_at(j), nulleq = sig / 10 * 10 == B2_SIG_NULLEQ_INT;
            if (an && bn) { if (nulleq) { r.nn[j] = 1; r.i[j] = 1; } continue; }
            if (an || bn) { if (nulleq) { r.nn[j] = 1; r.i[j] = 0; } continue; }
            const int c = cmp(key(args[0], j), key(args[1], j));
            bool res;
            switch (sig / 10 * 10) {
              case B2_SIG_LT_INT: res = c < 0; break; case B2_SIG_LE_INT: res = c <= 0; break;
              case B2_SIG_GT_INT: res = c > 0; break; case B2_SIG_GE_INT: res = c >= 0; break;
              case B2_SIG_NE_INT: res = c != 0; break; default: res = c == 0; break;
            }
            r.nn[j] = 1; r.i[j] = res;
          }
          st.push_back(std::move(r));
          continue;
        }
      }
      if (nd.sig == B2_SIG_IN_INT || nd.sig == B2_SIG_IN_REAL) {
        // compare_in_int_type_by_hash :217-258 / compare_in_by_hash :178-215 (varg): args[0] IN (args[1..])
        if ((int)st.size() < na || na < 1) { *err = Error::make(B2_ERR_INVALID_ARG, "bad rpn arity"); return false; }
        std::vector<Val> args(st.end() - na, st.end());
        st.resize(st.size() - na);
        Val r; r.et = ET_INT; r.is_unsigned = false; r.nn.assign(n, 0); r.i.assign(n, 0);
        for (size_t j = 0; j < n; ++j) {
          if (args[0].null_at(j)) continue;
          bool hit = false, default_null = false;
          for (int i = 1; i < na; ++i) {
            if (args[i].null_at(j)) { default_null = true; continue; }
            if (nd.sig == B2_SIG_IN_REAL) hit |= args[0].real_at(j) == args[i].real_at(j);
            else {
              int64_t base_val = args[0].int_at(j), v = args[i].int_at(j);
              hit |= base_val == v && (base_val >= 0 || args[0].is_unsigned == args[i].is_unsigned);
            }
          }
          if (hit) { r.nn[j] = 1; r.i[j] = 1; }
          else if (!default_null) { r.nn[j] = 1; r.i[j] = 0; }
        }
        st.push_back(std::move(r));
        continue;
      }
      if (nd.sig == B2_SIG_IF_INT || nd.sig == B2_SIG_IF_REAL || nd.sig == B2_SIG_CASE_WHEN_INT || nd.sig == B2_SIG_CASE_WHEN_REAL ||
          nd.sig == B2_SIG_COALESCE_INT || nd.sig == B2_SIG_COALESCE_REAL) {
        // if_condition impl_control.rs:88-100, case_when :34-50 (chunks of [cond, value], a trailing single value is the
        // ELSE), coalesce impl_compare.rs:239-248.  All arguments are already evaluated: only one is picked per row.
        const bool is_if = nd.sig == B2_SIG_IF_INT || nd.sig == B2_SIG_IF_REAL, is_co = nd.sig == B2_SIG_COALESCE_INT || nd.sig == B2_SIG_COALESCE_REAL;
        const bool real = nd.sig == B2_SIG_IF_REAL || nd.sig == B2_SIG_CASE_WHEN_REAL || nd.sig == B2_SIG_COALESCE_REAL;
        if ((int)st.size() < na || na < 1 || (is_if && na != 3)) { *err = Error::make(B2_ERR_INVALID_ARG, "bad rpn arity"); return false; }
        std::vector<Val> args(st.end() - na, st.end());
        st.resize(st.size() - na);
        Val r; r.et = real ? ET_REAL : ET_INT; r.is_unsigned = nd.field_flag & B2_FLAG_UNSIGNED; r.nn.assign(n, 0);
        if (real) r.f.assign(n, 0); else r.i.assign(n, 0);
        for (size_t j = 0; j < n; ++j) {
          int pick = -1;
          if (is_co) { for (int i = 0; i < na && pick < 0; ++i) if (!args[i].null_at(j)) pick = i; }
          else if (is_if) pick = (!args[0].null_at(j) && args[0].int_at(j) != 0) ? 1 : 2;
          else {
            for (int i = 0; i + 1 < na && pick < 0; i += 2) if (!args[i].null_at(j) && args[i].int_at(j) != 0) pick = i + 1;
            if (pick < 0 && (na & 1)) pick = na - 1;
          }
          if (pick < 0 || args[pick].null_at(j)) continue;
          r.nn[j] = 1;
          if (real) r.f[j] = args[pick].real_at(j); else r.i[j] = args[pick].int_at(j);
        }
        st.push_back(std::move(r));
        continue;
      }
      if ((int)st.size() < na || na < 1 || na > 2) { *err = Error::make(B2_ERR_INVALID_ARG, "bad rpn arity"); return false; }
      Val b; if (na == 2) { b = std::move(st.back()); st.pop_back(); }
      Val a = std::move(st.back()); st.pop_back();
      Val r; r.et = ET_INT; r.is_unsigned = nd.field_flag & B2_FLAG_UNSIGNED; r.nn.assign(n, 0); r.i.assign(n, 0);
      int sig = nd.sig;
      bool is_cmp_int = sig == B2_SIG_LT_INT || sig == B2_SIG_LE_INT || sig == B2_SIG_GT_INT || sig == B2_SIG_GE_INT || sig == B2_SIG_EQ_INT || sig == B2_SIG_NE_INT || sig == B2_SIG_NULLEQ_INT;
      bool is_cmp_real = sig == B2_SIG_LT_REAL || sig == B2_SIG_LE_REAL || sig == B2_SIG_GT_REAL || sig == B2_SIG_GE_REAL || sig == B2_SIG_EQ_REAL || sig == B2_SIG_NE_REAL || sig == B2_SIG_NULLEQ_REAL;
      bool is_arith_real = sig == B2_SIG_PLUS_REAL || sig == B2_SIG_MINUS_REAL || sig == B2_SIG_MULTIPLY_REAL || sig == B2_SIG_MOD_REAL || sig == B2_SIG_DIVIDE_REAL ||
                           sig == B2_SIG_IF_NULL_REAL || sig == B2_SIG_UNARY_MINUS_REAL || sig == B2_SIG_ABS_REAL || sig == B2_SIG_CAST_INT_AS_REAL ||
                           sig == B2_SIG_CAST_REAL_AS_REAL;
      if (is_arith_real) { r.et = ET_REAL; r.f.assign(n, 0); r.i.clear(); }
      for (size_t j = 0; j < n; ++j) {
        bool an = a.null_at(j), bn = na == 2 ? b.null_at(j) : false;
        if (is_cmp_int || is_cmp_real) {
          bool nulleq = sig == B2_SIG_NULLEQ_INT || sig == B2_SIG_NULLEQ_REAL;
          if (an && bn) { if (nulleq) { r.nn[j] = 1; r.i[j] = 1; } continue; }
          if (an || bn) { if (nulleq) { r.nn[j] = 1; r.i[j] = 0; } continue; }
          int c;
          if (is_cmp_int) c = cmp_int(a.int_at(j), a.is_unsigned, b.int_at(j), b.is_unsigned);
          else { double x = a.real_at(j), y = b.real_at(j); c = x < y ? -1 : x > y; }
          bool res;
          switch (sig - (is_cmp_real ? 1 : 0)) {
            case B2_SIG_LT_INT: res = c < 0; break; case B2_SIG_LE_INT: res = c <= 0; break;
            case B2_SIG_GT_INT: res = c > 0; break; case B2_SIG_GE_INT: res = c >= 0; break;
            case B2_SIG_EQ_INT: res = c == 0; break; case B2_SIG_NE_INT: res = c != 0; break;
            default: res = c == 0; break;
          }
          r.nn[j] = 1; r.i[j] = res;
          continue;
        }
        switch (sig) {
          case B2_SIG_LOGICAL_AND:  // impl_op.rs:8-16
            if ((!an && a.int_at(j) == 0) || (!bn && b.int_at(j) == 0)) { r.nn[j] = 1; r.i[j] = 0; }
            else if (an || bn) {} else { r.nn[j] = 1; r.i[j] = 1; }
            break;
          case B2_SIG_LOGICAL_OR:   // :18-28
            if (!an && !bn && a.int_at(j) == 0 && b.int_at(j) == 0) { r.nn[j] = 1; r.i[j] = 0; }
            else if ((an && bn) || (an && b.int_at(j) == 0) || (bn && a.int_at(j) == 0)) {}
            else { r.nn[j] = 1; r.i[j] = 1; }
            break;
          case B2_SIG_LOGICAL_XOR:  // :30-39
            if (!an && !bn) { r.nn[j] = 1; r.i[j] = (a.int_at(j) == 0) ^ (b.int_at(j) == 0); }
            break;
          case B2_SIG_UNARY_NOT_INT: if (!an) { r.nn[j] = 1; r.i[j] = a.int_at(j) == 0; } break;
          case B2_SIG_UNARY_NOT_REAL: if (!an) { r.nn[j] = 1; r.i[j] = a.real_at(j) == 0.0; } break;
          case B2_SIG_INT_IS_NULL: case B2_SIG_REAL_IS_NULL: r.nn[j] = 1; r.i[j] = an; break;
          case B2_SIG_INT_IS_TRUE: r.nn[j] = 1; r.i[j] = !an && a.int_at(j) != 0; break;
          case B2_SIG_REAL_IS_TRUE: r.nn[j] = 1; r.i[j] = !an && a.real_at(j) != 0.0; break;
          case B2_SIG_INT_IS_FALSE: r.nn[j] = 1; r.i[j] = !an && a.int_at(j) == 0; break;
          case B2_SIG_REAL_IS_FALSE: r.nn[j] = 1; r.i[j] = !an && a.real_at(j) == 0.0; break;
          case B2_SIG_PLUS_INT: case B2_SIG_MINUS_INT: case B2_SIG_MULTIPLY_INT: case B2_SIG_MULTIPLY_INT_UNSIGNED: {
            if (an || bn) break;
            int64_t x = a.int_at(j), y = b.int_at(j), z = 0;
            bool xu = a.is_unsigned, yu = b.is_unsigned, ovf = false;
            if (sig == B2_SIG_MULTIPLY_INT_UNSIGNED) xu = yu = true;
            if (sig == B2_SIG_PLUS_INT) {  // impl_arithmetic.rs:42-96
              if (!xu && !yu) ovf = __builtin_add_overflow(x, y, &z);
              else if (xu && yu) { uint64_t w; ovf = __builtin_add_overflow((uint64_t)x, (uint64_t)y, &w); z = (int64_t)w; }
              else {
                int64_t s = xu ? y : x; uint64_t u = (uint64_t)(xu ? x : y), w;
                if (s >= 0) ovf = __builtin_add_overflow((uint64_t)s, u, &w); else ovf = __builtin_sub_overflow(u, (uint64_t)0 - (uint64_t)s, &w);
                z = (int64_t)w;
              }
            } else if (sig == B2_SIG_MINUS_INT) {  // :126-190
              if (!xu && !yu) ovf = __builtin_sub_overflow(x, y, &z);
              else if (xu && yu) { uint64_t w; ovf = __builtin_sub_overflow((uint64_t)x, (uint64_t)y, &w); z = (int64_t)w; }
              else if (!xu && yu) { uint64_t w = 0; if (x >= 0) ovf = __builtin_sub_overflow((uint64_t)x, (uint64_t)y, &w); else ovf = true; z = (int64_t)w; }
              else { uint64_t w; if (y >= 0) ovf = __builtin_sub_overflow((uint64_t)x, (uint64_t)y, &w); else ovf = __builtin_add_overflow((uint64_t)x, (uint64_t)0 - (uint64_t)y, &w); z = (int64_t)w; }
            } else {  // multiply :345-398
              if (!xu && !yu) ovf = __builtin_mul_overflow(x, y, &z);
              else if (xu && yu) { uint64_t w; ovf = __builtin_mul_overflow((uint64_t)x, (uint64_t)y, &w); z = (int64_t)w; }
              else { int64_t s = xu ? y : x; uint64_t u = (uint64_t)(xu ? x : y), w = 0; if (s >= 0) ovf = __builtin_mul_overflow((uint64_t)s, u, &w); else ovf = true; z = (int64_t)w; }
            }
            if (ovf) { *err = overflow_err((xu || yu) ? "BIGINT UNSIGNED" : "BIGINT"); return false; }
            r.nn[j] = 1; r.i[j] = z;
            break;
          }
          case B2_SIG_PLUS_REAL: case B2_SIG_MINUS_REAL: case B2_SIG_MULTIPLY_REAL: {
            if (an || bn) break;
            double x = a.real_at(j), y = b.real_at(j);
            double z = sig == B2_SIG_PLUS_REAL ? x + y : sig == B2_SIG_MINUS_REAL ? x - y : x * y;
            bool bad = sig == B2_SIG_MULTIPLY_REAL ? std::isinf(z) : !std::isfinite(z);
            if (bad) { *err = overflow_err("DOUBLE"); return false; }
            r.nn[j] = 1; r.f[j] = z;
            break;
          }
          case B2_SIG_IF_NULL_INT:  // impl_control.rs:7-14
            if (!an) { r.nn[j] = 1; r.i[j] = a.int_at(j); } else if (!bn) { r.nn[j] = 1; r.i[j] = b.int_at(j); }
            break;
          case B2_SIG_IF_NULL_REAL:
            if (!an) { r.nn[j] = 1; r.f[j] = a.real_at(j); } else if (!bn) { r.nn[j] = 1; r.f[j] = b.real_at(j); }
            break;
          case B2_SIG_UNARY_MINUS_INT: {  // impl_op.rs:70-101; lib.rs map_unary_minus_int_func: unsigned argument -> unary_minus_uint
            if (an) break;
            int64_t x = a.int_at(j);
            if (a.is_unsigned ? (uint64_t)x > (uint64_t)INT64_MAX + 1 : x == INT64_MIN) { *err = overflow_err("BIGINT"); return false; }
            r.nn[j] = 1; r.i[j] = (int64_t)(0 - (uint64_t)x);
            break;
          }
          case B2_SIG_UNARY_MINUS_REAL: if (!an) { r.nn[j] = 1; r.f[j] = -a.real_at(j); } break;  // :103-107
          case B2_SIG_ABS_INT: {  // impl_math.rs:224-231
            if (an) break;
            int64_t x = a.int_at(j);
            if (x == INT64_MIN) { *err = overflow_err("BIGINT"); return false; }
            r.nn[j] = 1; r.i[j] = x < 0 ? -x : x;
            break;
          }
          case B2_SIG_ABS_UINT: if (!an) { r.nn[j] = 1; r.i[j] = a.int_at(j); } break;  // :233-237
          case B2_SIG_BIT_AND: if (!an && !bn) { r.nn[j] = 1; r.i[j] = a.int_at(j) & b.int_at(j); } break;  // impl_op.rs:144-175
          case B2_SIG_BIT_OR: if (!an && !bn) { r.nn[j] = 1; r.i[j] = a.int_at(j) | b.int_at(j); } break;
          case B2_SIG_BIT_XOR: if (!an && !bn) { r.nn[j] = 1; r.i[j] = a.int_at(j) ^ b.int_at(j); } break;
          case B2_SIG_BIT_NEG: if (!an) { r.nn[j] = 1; r.i[j] = ~a.int_at(j); } break;
          // impl_cast.rs:281-305 (cast_signed_int_as_unsigned_int / cast_int_as_int_others: the bits stay; in_union is false),
          // :466-501 (signed -> signed real: `as f64`; any unsigned side: `as u64 as f64`), :505-507
          case B2_SIG_CAST_INT_AS_INT: if (!an) { r.nn[j] = 1; r.i[j] = a.int_at(j); } break;
          case B2_SIG_CAST_INT_AS_REAL:
            if (!an) { r.nn[j] = 1; r.f[j] = (a.is_unsigned || (nd.field_flag & B2_FLAG_UNSIGNED)) ? (double)(uint64_t)a.int_at(j) : (double)a.int_at(j); }
            break;
          case B2_SIG_CAST_REAL_AS_REAL: if (!an) { r.nn[j] = 1; r.f[j] = a.real_at(j); } break;
          case B2_SIG_ABS_REAL: if (!an) { r.nn[j] = 1; r.f[j] = std::fabs(a.real_at(j)); } break;  // :239-243
          case B2_SIG_INT_DIVIDE_INT: {  // impl_arithmetic.rs:396-455; helpers codec/overflow.rs:9-58
            if (an || bn) break;
            int64_t x = a.int_at(j), y = b.int_at(j);
            if (y == 0) break;  // Ok(None)
            bool ovf = false; int64_t z = 0;
            if (!a.is_unsigned && !b.is_unsigned) { if (x == INT64_MIN && y == -1) ovf = true; else z = x / y; }                       // div_i64
            else if (!a.is_unsigned && b.is_unsigned) { if (x < 0) ovf = (0 - (uint64_t)x) >= (uint64_t)y; else z = (int64_t)((uint64_t)x / (uint64_t)y); }  // div_i64_with_u64
            else if (a.is_unsigned && b.is_unsigned) z = (int64_t)((uint64_t)x / (uint64_t)y);
            else { if (y < 0) ovf = x != 0 && (0 - (uint64_t)y) <= (uint64_t)x; else z = (int64_t)((uint64_t)x / (uint64_t)y); }   // div_u64_with_i64
            if (ovf) { *err = overflow_err("UNSIGNED BIGINT"); return false; }
            r.nn[j] = 1; r.i[j] = z;
            break;
          }
          case B2_SIG_MOD_INT: {  // impl_arithmetic.rs:215-278
            if (an || bn) break;
            int64_t x = a.int_at(j), y = b.int_at(j);
            if (y == 0) break;
            uint64_t ax = x < 0 ? 0 - (uint64_t)x : (uint64_t)x, ay = y < 0 ? 0 - (uint64_t)y : (uint64_t)y;  // overflowing_abs as u64
            int64_t z;
            if (!a.is_unsigned && !b.is_unsigned) z = y == -1 ? 0 : x % y;  // (i64::MIN % -1 panics in the reference; 0 is the value)
            else if (!a.is_unsigned && b.is_unsigned) z = x > 0 ? (int64_t)((uint64_t)x % (uint64_t)y) : (int64_t)(0 - ax % (uint64_t)y);
            else if (a.is_unsigned && !b.is_unsigned) z = (int64_t)((uint64_t)x % ay);
            else z = (int64_t)((uint64_t)x % (uint64_t)y);
            r.nn[j] = 1; r.i[j] = z;
            break;
          }
          case B2_SIG_DIVIDE_REAL: {  // impl_arithmetic.rs:515-533
            if (an || bn) break;
            if (b.real_at(j) == 0.0) { warning_count() += 1; break; }  // ctx.handle_division_by_zero(): a warning, the result is NULL
            double z = a.real_at(j) / b.real_at(j);
            if (std::isinf(z)) { *err = overflow_err("DOUBLE"); return false; }
            r.nn[j] = 1; r.f[j] = z;
            break;
          }
          case B2_SIG_MOD_REAL: {  // :280-291
            if (an || bn) break;
            if (b.real_at(j) == 0.0) break;
            r.nn[j] = 1; r.f[j] = std::fmod(a.real_at(j), b.real_at(j));
            break;
          }
          default: *err = Error::make(B2_ERR_UNSUPPORTED, "scalar function sig " + std::to_string(sig)); return false;
        }
      }
      st.push_back(std::move(r));
    } else { *err = Error::make(B2_ERR_INVALID_ARG, "bad rpn node kind"); return false; }
  }
  if (st.size() != 1) { *err = Error::make(B2_ERR_INVALID_ARG, "rpn does not reduce to one value"); return false; }
  *result = std::move(st.back());
  return true;
}

// ---- selection ----
struct SelectionExecutor : Executor {
  std::unique_ptr<Executor> src;
  std::vector<b2_rpn_expr> conditions;
  const std::vector<FieldType>& schema() const override { return src->schema(); }
  ForwardScanner* scanner() override { return src->scanner(); }
  void next_batch(size_t scan_rows, Batch* out) override {  // :81-139, 162-195
    src->next_batch(scan_rows, out);
    for (auto& cond : conditions) {
      if (out->logical_rows.empty()) break;
      ExprCtx cx{&src->schema(), out};
      Val v; Error e;
      if (!rpn_eval(cond, cx, &v, &e)) { out->logical_rows.clear(); out->err = e; return; }  // :225-236 discard batch
      std::vector<size_t> kept;
      for (size_t j = 0; j < out->logical_rows.size(); ++j) {
        bool t = !v.null_at(j) && (v.et == ET_REAL ? v.real_at(j) != 0.0 : v.int_at(j) != 0);  // AsMySqlBool data_type/mod.rs:58-110
        if (t) kept.push_back(out->logical_rows[j]);
      }
      out->logical_rows.swap(kept);
    }
  }
};

// ---- projection: projection_executor.rs:171-240 ----
struct ProjectionExecutor : Executor {
  std::unique_ptr<Executor> src;
  std::vector<b2_rpn_expr> exprs;
  std::vector<FieldType> schema_;
  const std::vector<FieldType>& schema() const override { return schema_; }
  ForwardScanner* scanner() override { return src->scanner(); }
  void next_batch(size_t scan_rows, Batch* out) override {
    Batch b;
    src->next_batch(scan_rows, &b);
    out->is_drained = b.is_drained; out->err = b.err;
    out->cols.clear(); out->logical_rows.clear();
    size_t n = b.logical_rows.size();
    if (!b.err.ok() || n == 0) { out->cols.assign(exprs.size(), LazyColumn()); for (auto& c : out->cols) c.decoded = true; return; }
    ExprCtx cx{&src->schema(), &b};
    for (size_t k = 0; k < exprs.size(); ++k) {
      Val v; Error e;
      if (!rpn_eval(exprs[k], cx, &v, &e)) {  // :207-211: the error ends the request, the batch's rows are dropped
        out->err = e; out->cols.assign(exprs.size(), LazyColumn()); for (auto& c : out->cols) c.decoded = true;
        return;
      }
      LazyColumn c; c.decoded = true; c.et = v.et == ET_REAL ? ET_REAL : ET_INT;
      for (size_t j = 0; j < n; ++j) {
        bool nul = v.null_at(j);
        c.nn.push_back(!nul);
        if (c.et == ET_REAL) c.f64.push_back(nul ? 0 : v.real_at(j)); else c.i64.push_back(nul ? 0 : v.int_at(j));
      }
      out->cols.push_back(std::move(c));
    }
    for (size_t j = 0; j < n; ++j) out->logical_rows.push_back(j);
  }
};

// ---- limit: limit_executor.rs:11-80 ----
struct LimitExecutor : Executor {
  std::unique_ptr<Executor> src;
  size_t remaining_rows = 0;
  bool is_src_scan_executor = false;
  const std::vector<FieldType>& schema() const override { return src->schema(); }
  ForwardScanner* scanner() override { return src->scanner(); }
  void next_batch(size_t scan_rows, Batch* out) override {
    size_t real_scan_rows = is_src_scan_executor ? std::min(scan_rows, remaining_rows) : scan_rows;  // :56-60
    src->next_batch(real_scan_rows, out);
    if (out->logical_rows.size() < remaining_rows) remaining_rows -= out->logical_rows.size();
    else { out->logical_rows.resize(remaining_rows); out->is_drained = true; remaining_rows = 0; }  // :70-77
  }
};

// ---- aggregation ----
struct AggState {
  uint64_t count = 0;
  Decimal dsum = dec_zero(); double fsum = 0; bool has_value = false;
  int64_t ext_i = 0; double ext_f = 0;  // MAX / MIN extremum (impl_max_min.rs:425-560), valid when has_value
};
struct AggFn { int kind; b2_rpn_expr arg; EvalType arg_et; bool arg_unsigned; };

struct AggExecutor : Executor {  // simple (no group by) or fast-hash (one group-by expr)
  std::unique_ptr<Executor> src;
  std::vector<AggFn> fns;
  bool has_group = false; b2_rpn_expr group_by; FieldType group_ft;
  std::vector<FieldType> schema_;
  // groups in first-seen order; key (is_null, bits)
  struct KeyHash { size_t operator()(const std::pair<bool, int64_t>& k) const { return std::hash<int64_t>()(k.second) ^ (k.first ? 0x9e3779b97f4a7c15ull : 0); } };
  std::unordered_map<std::pair<bool, int64_t>, size_t, KeyHash> groups;
  std::vector<std::pair<bool, int64_t>> group_keys;
  std::vector<AggState> states;  // group * fns
  bool any_input = false, done = false;
  EvalType group_et = ET_INT;
  // BatchSlowHashAggregation (slow_hash_aggr_executor.rs:209-420): two or more group-by expressions.  The reference keys
  // its map on the concatenated encode_sort_key bytes of the values; for Int / Real columns that is the value bits plus
  // NULL-ness per column (no -0.0 / 0.0 folding: the two encode to different bytes).
  std::vector<b2_rpn_expr> multi_by; std::vector<FieldType> multi_ft;
  std::map<std::vector<std::pair<bool, int64_t>>, size_t> multi_groups;
  std::vector<std::vector<std::pair<bool, int64_t>>> multi_keys;

  const std::vector<FieldType>& schema() const override { return schema_; }
  ForwardScanner* scanner() override { return src->scanner(); }

  bool update(AggState& s, const AggFn& f, const Val& v, size_t j, Error* err) {
    bool isnull = v.null_at(j);
    switch (f.kind) {
      case B2_AGG_COUNT: if (!isnull) s.count++; return true;
      case B2_AGG_AVG: if (!isnull) s.count++;  // fallthrough: impl_avg.rs:120-131 count + sum
      case B2_AGG_SUM:
        if (isnull) return true;
        if (v.et == ET_REAL) s.fsum += v.real_at(j);
        else {
          Decimal d = v.is_unsigned ? dec_from_u64((uint64_t)v.int_at(j)) : dec_from_i64(v.int_at(j));  // util.rs:46-52 cast
          Decimal r;
          DecRes st = dec_add(s.dsum, d, &r);
          if (st != DEC_OK) { *err = Error::make(B2_ERR_EVALUATE, st == DEC_OVERFLOW ? "DECIMAL value is out of range" : "Data truncated", st == DEC_OVERFLOW ? B2_MYSQL_ERR_DATA_OUT_OF_RANGE : B2_MYSQL_ERR_TRUNCATED); return false; }
          s.dsum = r;
        }
        s.has_value = true;
        return true;
      case B2_AGG_MAX: case B2_AGG_MIN: {
        // AggFnStateExtremumForInt::update_concrete :517-541 / AggFnStateExtremum :438-452: the first non-NULL value,
        // then replaced only when the stored extremum compares on the wrong side (E::ORD: Less for MAX, Greater for MIN)
        if (isnull) return true;
        const bool is_max = f.kind == B2_AGG_MAX;
        if (v.et == ET_REAL) {
          double x = v.real_at(j);
          if (!s.has_value || (is_max ? s.ext_f < x : s.ext_f > x)) s.ext_f = x;
        } else {
          int64_t x = v.int_at(j);
          bool replace = !s.has_value;
          if (!replace) replace = f.arg_unsigned ? (is_max ? (uint64_t)s.ext_i < (uint64_t)x : (uint64_t)s.ext_i > (uint64_t)x) : (is_max ? s.ext_i < x : s.ext_i > x);
          if (replace) s.ext_i = x;
        }
        s.has_value = true;
        return true;
      }
      default: *err = Error::make(B2_ERR_UNSUPPORTED, "aggregate kind"); return false;
    }
  }

  void next_batch(size_t, Batch* out) override {  // aggr_executor.rs:206-303: drains source with BATCH_MAX_SIZE pulls
    out->cols.clear(); out->logical_rows.clear(); out->is_drained = true; out->err = Error();
    if (done) return;
    done = true;
    for (;;) {
      Batch b;
      src->next_batch(BATCH_MAX_SIZE, &b);
      size_t n = b.logical_rows.size();
      if (n > 0) {
        any_input = true;
        ExprCtx cx{&src->schema(), &b};
        std::vector<size_t> row_group(n, 0);
        if (has_group) {  // calc_groups_each_row fast_hash_aggr_executor.rs:423-457
          Val g; Error e;
          if (!rpn_eval(group_by, cx, &g, &e)) { out->err = e; return; }
          for (size_t j = 0; j < n; ++j) {
            bool isnull = g.null_at(j);
            int64_t bits = 0;
            if (!isnull) { if (g.et == ET_REAL) { double d = g.real_at(j); if (d == 0.0) d = 0.0; memcpy(&bits, &d, 8); } else bits = g.int_at(j); }
            auto key = std::make_pair(isnull, bits);
            auto it = groups.find(key);
            if (it == groups.end()) { size_t gi = group_keys.size(); groups.emplace(key, gi); group_keys.push_back(key); states.resize(states.size() + fns.size()); row_group[j] = gi; }
            else row_group[j] = it->second;
          }
        } else if (!multi_by.empty()) {
          std::vector<Val> gv(multi_by.size());
          for (size_t q = 0; q < multi_by.size(); ++q) { Error e; if (!rpn_eval(multi_by[q], cx, &gv[q], &e)) { out->err = e; return; } }
          for (size_t j = 0; j < n; ++j) {
            std::vector<std::pair<bool, int64_t>> key;
            for (size_t q = 0; q < multi_by.size(); ++q) {
              bool isnull = gv[q].null_at(j);
              int64_t bits = 0;
              if (!isnull) { if (gv[q].et == ET_REAL) { double d = gv[q].real_at(j); memcpy(&bits, &d, 8); } else bits = gv[q].int_at(j); }
              key.emplace_back(isnull, bits);
            }
            auto it = multi_groups.find(key);
            if (it == multi_groups.end()) { size_t gi = multi_keys.size(); multi_groups.emplace(key, gi); multi_keys.push_back(key); states.resize(states.size() + fns.size()); row_group[j] = gi; }
            else row_group[j] = it->second;
          }
        } else if (states.empty()) states.resize(fns.size());
        for (size_t fi = 0; fi < fns.size(); ++fi) {
          Val v; Error e;
          if (!rpn_eval(fns[fi].arg, cx, &v, &e)) { out->err = e; return; }
          for (size_t j = 0; j < n; ++j) {
            if (!update(states[row_group[j] * fns.size() + fi], fns[fi], v, j, &e)) { out->err = e; return; }
          }
        }
      }
      if (!b.err.ok()) { out->err = b.err; return; }
      if (b.is_drained) break;
    }
    // emit: aggregate result columns then the group-by column (fast_hash_aggr_executor.rs:383-413)
    size_t ngroups = has_group ? group_keys.size() : (!multi_by.empty() ? multi_keys.size() : (any_input ? 1 : 0));  // simple_aggr_executor.rs:141-148
    if (!has_group && multi_by.empty() && any_input && states.empty()) states.resize(fns.size());
    for (size_t fi = 0; fi < fns.size(); ++fi) {
      const AggFn& f = fns[fi];
      if (f.kind == B2_AGG_COUNT || f.kind == B2_AGG_AVG) {
        LazyColumn c; c.decoded = true; c.et = ET_INT;
        for (size_t g = 0; g < ngroups; ++g) c.push_int(true, (int64_t)states[g * fns.size() + fi].count);
        out->cols.push_back(std::move(c));
      }
      if (f.kind == B2_AGG_SUM || f.kind == B2_AGG_AVG) {
        LazyColumn c; c.decoded = true;
        if (f.arg_et == ET_REAL) { c.et = ET_REAL; for (size_t g = 0; g < ngroups; ++g) { const AggState& s = states[g * fns.size() + fi]; c.f64.push_back(s.has_value ? s.fsum : 0); c.nn.push_back(s.has_value); } }
        else { c.et = ET_DECIMAL; for (size_t g = 0; g < ngroups; ++g) { const AggState& s = states[g * fns.size() + fi]; dec_col.push_back(s.dsum); c.i64.push_back((int64_t)dec_col.size() - 1); c.nn.push_back(s.has_value); } }
        out->cols.push_back(std::move(c));
      }
      if (f.kind == B2_AGG_MAX || f.kind == B2_AGG_MIN) {  // push_result :471-474 / :553-556: Option<T>
        LazyColumn c; c.decoded = true; c.et = f.arg_et == ET_REAL ? ET_REAL : ET_INT;
        for (size_t g = 0; g < ngroups; ++g) {
          const AggState& s = states[g * fns.size() + fi];
          if (f.arg_et == ET_REAL) { c.f64.push_back(s.has_value ? s.ext_f : 0); c.nn.push_back(s.has_value); }
          else c.push_int(s.has_value, s.ext_i);
        }
        out->cols.push_back(std::move(c));
      }
    }
    if (has_group) {
      LazyColumn c; c.decoded = true; c.et = group_et;
      for (auto& k : group_keys) {
        if (group_et == ET_REAL) { double d; memcpy(&d, &k.second, 8); c.f64.push_back(k.first ? 0 : d); c.nn.push_back(!k.first); }
        else c.push_int(!k.first, k.second);
      }
      out->cols.push_back(std::move(c));
    }
    for (size_t q = 0; q < multi_by.size(); ++q) {  // the group-by columns, in order (slow_hash_aggr_executor.rs:388-420)
      LazyColumn c; c.decoded = true; c.et = eval_type_of(multi_ft[q].tp) == ET_REAL ? ET_REAL : ET_INT;
      for (auto& k : multi_keys) {
        if (c.et == ET_REAL) { double d; memcpy(&d, &k[q].second, 8); c.f64.push_back(k[q].first ? 0 : d); c.nn.push_back(!k[q].first); }
        else c.push_int(!k[q].first, k[q].second);
      }
      out->cols.push_back(std::move(c));
    }
    out->logical_rows.resize(ngroups);
    for (size_t i = 0; i < ngroups; ++i) out->logical_rows[i] = i;
  }
  std::vector<Decimal> dec_col;  // decimal cells referenced by index from ET_DECIMAL columns
};

// ---- TopN ----
struct TopNExecutor : Executor {
  std::unique_ptr<Executor> src;
  struct Ord { b2_rpn_expr expr; bool desc; };
  std::vector<Ord> order;
  uint64_t n = 0;
  bool done = false;
  struct Item {
    std::vector<uint8_t> knull; std::vector<int64_t> kbits; std::vector<uint8_t> kreal, kuns;
    std::vector<int64_t> ci; std::vector<double> cf; std::vector<uint8_t> cnn;  // decoded source row
  };
  const std::vector<FieldType>& schema() const override { return src->schema(); }
  ForwardScanner* scanner() override { return src->scanner(); }

  // HeapItemUnsafe::cmp_sort_key top_n_heap.rs:188-222 + scalar.rs:374-411 (None < Some; unsigned compare by field type)
  int cmp(const Item& a, const Item& b) const {
    for (size_t k = 0; k < order.size(); ++k) {
      int c;
      if (a.knull[k] || b.knull[k]) c = (a.knull[k] && b.knull[k]) ? 0 : (a.knull[k] ? -1 : 1);
      else if (a.kreal[k]) { double x, y; memcpy(&x, &a.kbits[k], 8); memcpy(&y, &b.kbits[k], 8); c = x < y ? -1 : x > y; }
      else if (a.kuns[k]) c = (uint64_t)a.kbits[k] < (uint64_t)b.kbits[k] ? -1 : (uint64_t)a.kbits[k] > (uint64_t)b.kbits[k];
      else c = a.kbits[k] < b.kbits[k] ? -1 : a.kbits[k] > b.kbits[k];
      if (c == 0) continue;
      return order[k].desc ? -c : c;
    }
    return 0;
  }

  void next_batch(size_t, Batch* out) override {
    const auto& sch = src->schema();
    out->cols.clear(); out->logical_rows.clear(); out->is_drained = true; out->err = Error();
    if (done) return;
    done = true;
    if (n == 0) return;  // top_n_executor.rs:304-312
    std::vector<Item> heap;  // max-heap by cmp
    auto less = [this](const Item& a, const Item& b) { return cmp(a, b) < 0; };
    for (;;) {
      Batch b;
      src->next_batch(BATCH_MAX_SIZE, &b);
      size_t rows = b.logical_rows.size();
      if (rows > 0) {
        ExprCtx cx{&sch, &b};
        std::vector<Val> keys(order.size());
        for (size_t k = 0; k < order.size(); ++k) { Error e; if (!rpn_eval(order[k].expr, cx, &keys[k], &e)) { out->err = e; return; } }
        for (size_t ci = 0; ci < sch.size(); ++ci) { std::string perr; if (!ensure_decoded(b.cols[ci], sch[ci], b.logical_rows, &perr)) { out->err = Error::make(B2_ERR_CORRUPTED, perr); return; } }
        for (size_t j = 0; j < rows; ++j) {  // process_batch_input :204-273 -> add_row top_n_heap.rs:36-53
          Item it;
          for (size_t k = 0; k < order.size(); ++k) {
            const Val& v = keys[k];
            bool isnull = v.null_at(j);
            it.knull.push_back(isnull); it.kreal.push_back(v.et == ET_REAL); it.kuns.push_back(v.is_unsigned);
            int64_t bits = 0;
            if (!isnull) { if (v.et == ET_REAL) { double d = v.real_at(j); memcpy(&bits, &d, 8); } else bits = v.int_at(j); }
            it.kbits.push_back(bits);
          }
          if (heap.size() >= n && !(cmp(it, heap.front()) < 0)) continue;
          size_t r = b.logical_rows[j];
          for (size_t ci = 0; ci < sch.size(); ++ci) {
            const LazyColumn& c = b.cols[ci];
            it.cnn.push_back(c.nn[r]); it.ci.push_back(c.et == ET_INT ? c.i64[r] : 0); it.cf.push_back(c.et == ET_REAL ? c.f64[r] : 0);
          }
          if (heap.size() < n) { heap.push_back(std::move(it)); std::push_heap(heap.begin(), heap.end(), less); }
          else { std::pop_heap(heap.begin(), heap.end(), less); heap.back() = std::move(it); std::push_heap(heap.begin(), heap.end(), less); }
        }
      }
      if (!b.err.ok()) { out->err = b.err; return; }
      if (b.is_drained) break;
    }
    std::sort_heap(heap.begin(), heap.end(), less);
    out->cols.assign(sch.size(), LazyColumn());
    for (size_t ci = 0; ci < sch.size(); ++ci) { out->cols[ci].decoded = true; out->cols[ci].et = eval_type_of(sch[ci].tp); }
    for (auto& it : heap)
      for (size_t ci = 0; ci < sch.size(); ++ci) {
        LazyColumn& c = out->cols[ci];
        c.nn.push_back(it.cnn[ci]);
        if (c.et == ET_INT) c.i64.push_back(it.ci[ci]); else c.f64.push_back(it.cf[ci]);
      }
    out->logical_rows.resize(heap.size());
    for (size_t i = 0; i < heap.size(); ++i) out->logical_rows[i] = i;
  }
};

}  // namespace orc
