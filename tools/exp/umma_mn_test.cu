// Standalone probe of tcgen05.mma kind::tf32 with MN-major operands (one 128x256x8 instruction).
// A[m][k] and B[n][k] are written to shared memory by plain stores in the canonical MN-major
// SWIZZLE_128B layout; D is read back from TMEM and compared with the exact product.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -I n2nmn_b200/csrc -o gpurun_out/umma_mn_test tools/exp/umma_mn_test.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace n2nmn;

constexpr int M = 128, N = 256, K = 8;
// byte offset of element (mn, k) of an MN-major operand: 32-element blocks LBO apart, K rows of 128 B
__host__ __device__ inline uint32_t mn_offset(int mn, int k, uint32_t lbo) {
  uint32_t off = (mn / 32) * lbo + k * 128 + (mn % 32) * 4;
  // SWIZZLE_128B_BASE32B: the 32-byte chunk index (bits 5-6) is XORed with the row (bits 7-8)
  const uint32_t chunk = (off >> 5) & 3, row = (off >> 7) & 3;
  off = (off & ~0x60u) | ((chunk ^ row) << 5);
  return off;
}

__global__ void __launch_bounds__(128, 1) probe(const float* A, const float* B, float* D, uint32_t idesc_extra,
                                                uint32_t lbo, uint32_t sbo, int swap) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  uint8_t* sa = sm;
  uint8_t* sb = sm + 16384;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < M * K; i += blockDim.x) {
    const int m = i / K, k = i % K;
    *reinterpret_cast<float*>(sa + mn_offset(m, k, 4096)) = A[m * K + k];
  }
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i % K;
    *reinterpret_cast<float*>(sb + mn_offset(n, k, 4096)) = B[n * K + k];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (warp == 0) ptx::tmem_alloc<256>(&slot);
  ptx::fence_proxy_async();     // generic-proxy smem writes -> visible to the tensor core (async proxy)
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0 && ptx::elect_one()) {
    const uint32_t idesc = ptx::make_idesc_tf32(M, N) | idesc_extra;
    uint64_t da = ptx::make_smem_desc_sw128_mn(ptx::smem_u32(sa), swap ? sbo : lbo, swap ? lbo : sbo);
    uint64_t db = ptx::make_smem_desc_sw128_mn(ptx::smem_u32(sb), swap ? sbo : lbo, swap ? lbo : sbo);
    // layout type 1 = SWIZZLE_128B_BASE32B instead of 2
    da = (da & ~(7ull << 61)) | (1ull << 61);
    db = (db & ~(7ull << 61)) | (1ull << 61);
    ptx::umma_tf32(tmem, da, db, idesc, 0);
    ptx::umma_commit(&bar);
  }
  ptx::mbar_wait_bounded(&bar, 0);
  ptx::tc_fence_after();
  for (int cb = 0; cb < N; cb += 32) {
    float v[32];
    ptx::tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + cb, v);
    for (int c = 0; c < 32; ++c) D[(warp * 32 + lane) * N + cb + c] = v[c];
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc<256>(tmem); }
}

int main(int argc, char** argv) {
  const uint32_t extra = argc > 1 ? strtoul(argv[1], nullptr, 0) : ((1u << 15) | (1u << 16));
  const uint32_t lbo = argc > 2 ? atoi(argv[2]) : 4096, sbo = argc > 3 ? atoi(argv[3]) : 512;
  const int swap = argc > 4 ? atoi(argv[4]) : 0;
  float *hA = new float[M * K], *hB = new float[N * K], *hD = new float[M * N];
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) hA[m * K + k] = (float)((m * 7 + k * 3) % 11) - 5.f;
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) hB[n * K + k] = (float)((n * 5 + k) % 13) - 6.f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, M * K * 4); cudaMalloc(&dB, N * K * 4); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA, M * K * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, N * K * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, M * N * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768);
  probe<<<1, 128, 16384 + 32768>>>(dA, dB, dD, extra, lbo, sbo, swap);
  cudaError_t e = cudaDeviceSynchronize();
  printf("idesc_extra=%#x lbo=%u sbo=%u swap=%d -> %s\n", extra, lbo, sbo, swap, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
  int bad = 0, zero = 0;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    float ref = 0.f;
    for (int k = 0; k < K; ++k) ref += hA[m * K + k] * hB[n * K + k];
    if (hD[m * N + n] != ref) ++bad;
    if (hD[m * N + n] == 0.f) ++zero;
  }
  printf("mismatches %d of %d, zeros %d; D[0][0..3]=%g %g %g %g (ref %g) D[33][70]=%g\n", bad, M * N, zero,
         hD[0], hD[1], hD[2], hD[3], hA[0] * hB[0] + hA[1] * hB[1] + hA[2] * hB[2] + hA[3] * hB[3] + hA[4] * hB[4] +
         hA[5] * hB[5] + hA[6] * hB[6] + hA[7] * hB[7], hD[33 * N + 70]);
  return 0;
}
