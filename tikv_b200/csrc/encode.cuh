// Response encoders (encode.cu) shared with engine.cu.
#pragma once
#include <cuda_runtime.h>

#include "../../include/b2_copr.h"

namespace b2 {

// one decoded column of the batch being encoded (device pointers)
struct EncCol {
  const void* data;                  // i64 / f64 / b2_decimal cells
  const unsigned long long* bitmap;  // bit r = 1 -> non-null
  int kind;                          // B2_COL_*
  int is_f32;                        // field type FLOAT: 4-byte cells in TypeChunk
  int is_unsigned;                   // UINT_FLAG instead of INT_FLAG in TypeDefault
  unsigned int null_cnt;             // TypeChunk header
  unsigned long long chunk_off;      // byte offset of this column's block in the TypeChunk output
  const long long* offsets;          // B2_COL_BYTES / B2_COL_JSON: n_rows + 1 offsets into `data` (the byte heap)
  unsigned long long heap_len;       //   bytes of the heap in use (= offsets[n_rows])
};

cudaError_t launch_enc_null_count(const EncCol* cols, int n_cols, unsigned long long n_rows, unsigned int* counts, cudaStream_t s);
cudaError_t launch_enc_chunk(const EncCol* cols, int n_cols, unsigned long long n_rows, unsigned char* out, cudaStream_t s);
cudaError_t launch_enc_row_len(const EncCol* cols, int n_cols, unsigned long long n_rows, unsigned int* lens, cudaStream_t s);
size_t enc_scan_temp_bytes(unsigned long long n_rows);
cudaError_t launch_enc_scan(const unsigned int* lens, unsigned long long* offs, unsigned long long n_rows, void* temp, size_t temp_bytes, cudaStream_t s);
cudaError_t launch_enc_rows(const EncCol* cols, int n_cols, unsigned long long n_rows, const unsigned long long* row_offs, unsigned int fixed_len,
                            unsigned char* out, cudaStream_t s);

}  // namespace b2
