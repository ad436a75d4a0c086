// cp.async helpers (16-byte copies global -> shared, completion by commit groups) shared by the
// tile engines (mma_tile.cuh, backward.cuh).
#pragma once
#include "common.cuh"

namespace n2nmn {

static_assert(kTextRowsPerCta == 64, "the text kernels are written for 64-row groups");

// cp.async helpers (16-byte copies global -> shared, completion by commit groups)
__device__ __forceinline__ void tp_cp16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
               ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void tp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

}  // namespace n2nmn
