// Response encoding on the device: the rows of the batch an executor just produced, still resident in
// HBM as decoded columns, are written in the two wire formats of tipb::Chunk.rows_data
//   TypeChunk   one block per column: u32 len | u32 null_cnt | bitmap (if any NULL) | fixed-width cells
//               (tidb_query_datatype/src/codec/chunk/column.rs:50-70, 446-496, 1052-1072)
//   TypeDefault datum rows: per row, per column  NIL | INT/UINT flag + 8-byte memcomparable | FLOAT flag + 8 bytes |
//               DECIMAL flag + prec + frac + binary decimal
//               (lazy_column_vec.rs:172-187, vector.rs:362-470, datum_codec.rs:248-287, decimal.rs:2025-2132)
// so that only response bytes cross PCIe (runner.rs:1051-1088 does this on the CPU after every batch).
#include <cuda_runtime.h>

#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "encode.cuh"

namespace b2 {

__device__ __forceinline__ bool cell_non_null(const EncCol& c, unsigned long long r) { return (c.bitmap[r >> 6] >> (r & 63)) & 1ull; }

// NULL cells per column
__global__ void __launch_bounds__(256) enc_null_count_kernel(const EncCol* cols, unsigned long long n_rows, unsigned int* counts) {
  const EncCol c = cols[blockIdx.x];
  const unsigned long long words = (n_rows + 63) / 64;
  unsigned int cnt = 0;
  for (unsigned long long w = threadIdx.x; w < words; w += blockDim.x) {
    unsigned long long bits = ~c.bitmap[w];
    if (w == words - 1 && (n_rows & 63)) bits &= (1ull << (n_rows & 63)) - 1;
    cnt += __popcll(bits);
  }
  __shared__ unsigned int s[256];
  s[threadIdx.x] = cnt;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) s[threadIdx.x] += s[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[blockIdx.x] = s[0];
}

__device__ __forceinline__ void put_bytes_le(unsigned char* dst, unsigned long long v, int n) {
  if (n == 8 && ((unsigned long long)dst & 7) == 0) { *reinterpret_cast<unsigned long long*>(dst) = v; return; }
  if (n == 4 && ((unsigned long long)dst & 3) == 0) { *reinterpret_cast<unsigned int*>(dst) = (unsigned int)v; return; }
  for (int i = 0; i < n; ++i) dst[i] = (unsigned char)(v >> (8 * i));
}

// ---- TypeChunk: grid.y = column, grid.x strides over rows ----
__global__ void __launch_bounds__(256) enc_chunk_kernel(const EncCol* cols, unsigned long long n_rows, unsigned char* out) {
  const EncCol c = cols[blockIdx.y];
  unsigned char* base = out + c.chunk_off;
  const unsigned long long gtid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x, gsz = (unsigned long long)gridDim.x * blockDim.x;
  if (gtid == 0) { put_bytes_le(base, (unsigned int)n_rows, 4); put_bytes_le(base + 4, c.null_cnt, 4); }
  unsigned long long bm_bytes = 0;
  if (c.null_cnt) {
    bm_bytes = (n_rows + 7) / 8;
    for (unsigned long long j = gtid; j < bm_bytes; j += gsz) {
      unsigned int b = (unsigned int)(c.bitmap[j >> 3] >> ((j & 7) * 8)) & 0xffu;
      if (j == bm_bytes - 1 && (n_rows & 7)) b &= (1u << (n_rows & 7)) - 1;
      base[8 + j] = (unsigned char)b;
    }
  }
  unsigned char* data = base + 8 + bm_bytes;
  if (c.kind == B2_COL_BYTES || c.kind == B2_COL_JSON) {  // var-length column: (n + 1) i64 offsets, then the cells back to back
    for (unsigned long long i = gtid; i <= n_rows; i += gsz) put_bytes_le(data + i * 8, (unsigned long long)c.offsets[i], 8);
    unsigned char* heap = data + (n_rows + 1) * 8;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(c.data);
    for (unsigned long long i = gtid; i < c.heap_len; i += gsz) heap[i] = src[i];
  } else if (c.kind == B2_COL_DECIMAL) {
    const unsigned int* src = reinterpret_cast<const unsigned int*>(c.data);
    for (unsigned long long i = gtid; i < n_rows * 10; i += gsz) {  // 40-byte structs as 10 words; NULL cells are zero
      unsigned long long r = i / 10;
      put_bytes_le(data + i * 4, cell_non_null(c, r) ? src[i] : 0u, 4);
    }
  } else if (c.is_f32) {
    const double* src = reinterpret_cast<const double*>(c.data);
    for (unsigned long long i = gtid; i < n_rows; i += gsz) {
      float f = cell_non_null(c, i) ? (float)src[i] : 0.0f;
      put_bytes_le(data + i * 4, __float_as_uint(f), 4);
    }
  } else {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(c.data);
    for (unsigned long long i = gtid; i < n_rows; i += gsz) put_bytes_le(data + i * 8, cell_non_null(c, i) ? src[i] : 0ull, 8);
  }
}

// ---- binary decimal (decimal.rs:2025-2132 with (prec, frac) = prec_and_frac(), :1043-1051) ----
__device__ const unsigned int ENC_TEN_POW[10] = {1, 10, 100, 1000, 10000, 100000, 1000000, 10000000, 100000000, 1000000000};
__device__ const unsigned char ENC_DIG_2_BYTES[10] = {0, 1, 1, 2, 2, 3, 3, 4, 4, 4};

struct DecShape { int widx, int_cnt, prec; };
__device__ __forceinline__ DecShape dec_shape(const b2_decimal& d) {  // remove_leading_zeroes :1005-1018
  int cnt = d.int_cnt, i = ((cnt + 8) % 9) + 1, widx = 0;
  while (cnt > 0 && d.word_buf[widx] == 0) { cnt -= i; i = 9; widx++; }
  if (cnt > 0) {
    int k = (cnt - 1) % 9;
    while (ENC_TEN_POW[k] > d.word_buf[widx]) { k--; cnt--; }
  }
  DecShape s;
  s.widx = widx; s.int_cnt = cnt; s.prec = cnt + d.frac_cnt == 0 ? 1 : cnt + d.frac_cnt;
  return s;
}
__device__ __forceinline__ int dec_bin_len(const b2_decimal& d, const DecShape& s) {
  int int_cnt = s.prec - d.frac_cnt;
  return 2 + (int_cnt / 9) * 4 + ENC_DIG_2_BYTES[int_cnt % 9] + (d.frac_cnt / 9) * 4 + ENC_DIG_2_BYTES[d.frac_cnt % 9];
}
// writes prec, frac and the binary form; returns bytes written
__device__ int dec_write(unsigned char* o, const b2_decimal& d, const DecShape& s) {
  const int frac = d.frac_cnt;
  int n = 0, written = 0;
  o[n++] = (unsigned char)s.prec; o[n++] = (unsigned char)frac;
  unsigned int mask = d.negative ? 0xffffffffu : 0u;
  auto w_word = [&](unsigned int word, int size) {
    for (int i = 0; i < size; ++i) {
      unsigned char b = (unsigned char)(word >> (8 * (size - 1 - i)));
      if (written == 0 && i == 0) b ^= 0x80;
      o[n++] = b;
    }
    written += size;
  };
  int int_cnt = s.prec - frac;
  const int frac_words = frac / 9, trailing = frac % 9;
  const int src_frac_size = frac_words * 4 + ENC_DIG_2_BYTES[trailing];
  if (s.int_cnt + src_frac_size == 0) { mask = 0; int_cnt = 1; }
  const int int_size = (int_cnt / 9) * 4 + ENC_DIG_2_BYTES[int_cnt % 9];
  const int src_int_words = s.int_cnt / 9, src_leading = s.int_cnt % 9;
  const int src_int_size = src_int_words * 4 + ENC_DIG_2_BYTES[src_leading];
  for (int i = src_int_size; i < int_size; ++i) w_word(mask & 0xffu, 1);  // only the value 0 pads
  int widx = s.widx;
  if (src_leading > 0) { w_word((d.word_buf[widx] % ENC_TEN_POW[src_leading]) ^ mask, ENC_DIG_2_BYTES[src_leading]); widx++; }
  for (int k = 0; k < src_int_words + frac_words; ++k) w_word(d.word_buf[widx++] ^ mask, 4);
  if (trailing > 0) w_word((d.word_buf[widx] / ENC_TEN_POW[9 - trailing]) ^ mask, ENC_DIG_2_BYTES[trailing]);
  return n;
}

// ---- TypeDefault ----
__global__ void __launch_bounds__(256) enc_row_len_kernel(const EncCol* cols, int n_cols, unsigned long long n_rows, unsigned int* lens) {
  unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  unsigned int len = 0;
  for (int k = 0; k < n_cols; ++k) {
    const EncCol& c = cols[k];
    if (!cell_non_null(c, r)) len += 1;
    else if (c.kind == B2_COL_DECIMAL) {
      const b2_decimal d = reinterpret_cast<const b2_decimal*>(c.data)[r];
      len += 1 + dec_bin_len(d, dec_shape(d));
    } else len += 9;
  }
  lens[r] = len;
}

__device__ __forceinline__ void put_be64(unsigned char* o, unsigned long long v) {
  for (int i = 0; i < 8; ++i) o[i] = (unsigned char)(v >> (8 * (7 - i)));
}

// one thread per row; `row_offs` = exclusive scan of the row lengths, or nullptr when every row is `fixed_len` bytes
__global__ void __launch_bounds__(256) enc_rows_kernel(const EncCol* cols, int n_cols, unsigned long long n_rows, const unsigned long long* row_offs,
                                                       unsigned int fixed_len, unsigned char* out) {
  unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  unsigned char* o = out + (row_offs ? row_offs[r] : r * fixed_len);
  for (int k = 0; k < n_cols; ++k) {
    const EncCol& c = cols[k];
    if (!cell_non_null(c, r)) { *o++ = 0; continue; }  // NIL_FLAG
    if (c.kind == B2_COL_DECIMAL) {
      const b2_decimal d = reinterpret_cast<const b2_decimal*>(c.data)[r];
      *o++ = 6;  // DECIMAL_FLAG
      o += dec_write(o, d, dec_shape(d));
    } else if (c.kind == B2_COL_F64) {
      unsigned long long u = reinterpret_cast<const unsigned long long*>(c.data)[r];
      u = (u >> 63) ? ~u : (u | 0x8000000000000000ull);  // encode_f64 (tikv_util/src/codec/number.rs:27-34)
      *o++ = 5;  // FLOAT_FLAG
      put_be64(o, u); o += 8;
    } else {
      unsigned long long u = reinterpret_cast<const unsigned long long*>(c.data)[r];
      if (c.is_unsigned) { *o++ = 4; put_be64(o, u); }            // UINT_FLAG
      else { *o++ = 3; put_be64(o, u ^ 0x8000000000000000ull); }  // INT_FLAG, encode_i64
      o += 8;
    }
  }
}

struct U32ToU64 {
  __host__ __device__ unsigned long long operator()(const unsigned int& v) const { return (unsigned long long)v; }
};

cudaError_t launch_enc_null_count(const EncCol* cols, int n_cols, unsigned long long n_rows, unsigned int* counts, cudaStream_t s) {
  enc_null_count_kernel<<<n_cols, 256, 0, s>>>(cols, n_rows, counts);
  return cudaGetLastError();
}
cudaError_t launch_enc_chunk(const EncCol* cols, int n_cols, unsigned long long n_rows, unsigned char* out, cudaStream_t s) {
  unsigned long long work = n_rows * 10 / 256 + 1;
  dim3 grid((unsigned int)(work < 2048 ? work : 2048), (unsigned int)n_cols);
  enc_chunk_kernel<<<grid, 256, 0, s>>>(cols, n_rows, out);
  return cudaGetLastError();
}
cudaError_t launch_enc_row_len(const EncCol* cols, int n_cols, unsigned long long n_rows, unsigned int* lens, cudaStream_t s) {
  enc_row_len_kernel<<<(unsigned int)((n_rows + 255) / 256), 256, 0, s>>>(cols, n_cols, n_rows, lens);
  return cudaGetLastError();
}
size_t enc_scan_temp_bytes(unsigned long long n_rows) {
  size_t bytes = 0;
  cub::TransformInputIterator<unsigned long long, U32ToU64, const unsigned int*> it(nullptr, U32ToU64());
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, (unsigned long long*)nullptr, (int)n_rows);
  return bytes;
}
cudaError_t launch_enc_scan(const unsigned int* lens, unsigned long long* offs, unsigned long long n_rows, void* temp, size_t temp_bytes, cudaStream_t s) {
  cub::TransformInputIterator<unsigned long long, U32ToU64, const unsigned int*> it(lens, U32ToU64());
  return cub::DeviceScan::ExclusiveSum(temp, temp_bytes, it, offs, (int)n_rows, s);
}
cudaError_t launch_enc_rows(const EncCol* cols, int n_cols, unsigned long long n_rows, const unsigned long long* row_offs, unsigned int fixed_len,
                            unsigned char* out, cudaStream_t s) {
  enc_rows_kernel<<<(unsigned int)((n_rows + 255) / 256), 256, 0, s>>>(cols, n_cols, n_rows, row_offs, fixed_len, out);
  return cudaGetLastError();
}

}  // namespace b2
