// Run-time compiled, plan-specialised scan kernels (jit.cu).
#pragma once
#include <cuda_runtime.h>

#include <future>
#include <string>

#include "kernels.cuh"

namespace b2 {

struct JitKernel {
  bool ok = false;
  void* fn = nullptr;  // CUfunction
  mutable size_t max_dyn_smem = 48 * 1024;  // dynamic shared memory the function has been opted in to
  std::string error;
};

// NVRTC, the driver API and the kernel sources are all reachable from this process?
bool jit_available(std::string* why = nullptr);
// Starts (or joins) the compilation of the kernel specialised for `plan` on `device`; never blocks.
std::shared_future<JitKernel*> jit_get(int device, const DevPlan& plan);
// NVRTC only, no GPU: compile `plan`'s kernel into the on-disk cache (0 compiled, 1 already cached, < 0 failed)
int jit_precompile(const DevPlan& plan, std::string* error);
// process-wide: NVRTC compilations run, kernels served from the on-disk cache
void jit_counters(unsigned long long* nvrtc_compiles, unsigned long long* disk_hits);
int jit_max_blocks_per_sm(const JitKernel* k, size_t smem);
cudaError_t jit_launch(const JitKernel* k, const ScanArgs& a, int grid, size_t smem, cudaStream_t s);

}  // namespace b2
