// fc_eltwise of the answer heads with many classes (VQA: 3001) on the 5th-gen tensor cores with
// fp32 parity: scores[root, :] = ê[root, :]·W_out + b as THREE tcgen05.mma kind::tf32 products
//   ê_hi·W_hi + ê_lo·W_hi + ê_hi·W_lo      (x_hi = the fp32 value as the tensor core reads it, i.e.
//                                            truncated to TF32; x_lo = x - x_hi, exact in fp32)
// accumulated in one TMEM tile — the same error-compensated scheme as mma_tile.cuh, at the tcgen05
// issue rate instead of mma.sync's (~36 cycles per m16n8k8 per sub-partition).
// Operands, all K-major (channel contiguous), TMA SWIZZLE_128B boxes of 32 channels:
//   A_hi = ê [roots][Mp] exactly as head_kernel writes it, A_lo = its truncation remainder (second
//   buffer written by head_kernel), B_hi / B_lo = W_outᵀ [classes padded to 128][Mp] and its
//   remainder, prepared when the weights are set (out_wt_split_kernel).
// CTA = 128 roots x 128 classes, K streamed 32 channels per stage through a 3-stage ring of 64 KB;
// warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (12 instructions per stage),
// warps 2..5 = epilogue (TMEM -> + bias -> the roots' score rows).
#pragma once
#include "node_eval.cuh"
#include "ptx_sm100.cuh"

namespace n2nmn {

constexpr int kHtM = 128, kHtN = 128, kHtK = 32, kHtStages = 3, kHtThreads = 192;
constexpr int kHtTileBytes = 128 * kHtK * 4;            // 16 KB: 128 rows x 32 channels
constexpr int kHtStageBytes = 4 * kHtTileBytes;         // A_hi, A_lo, B_hi, B_lo
constexpr size_t kHtSmemBytes = (size_t)kHtStages * kHtStageBytes + 256;

struct HeadTailMaps { CUtensorMap a_hi, a_lo, b_hi, b_lo; };   // box (32 channels, 128 rows)

// W_out [M][C] -> W_outᵀ [Cpad][Mp] as (hi = the value, lo = value - trunc_tf32(value)), zero padded.
__global__ void out_wt_split_kernel(const float* __restrict__ W, int M, int C, float* __restrict__ hi,
                                    float* __restrict__ lo, int Mp, int Cpad) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, cc = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < M && cc < C) ? W[(size_t)k * C + cc] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int cc = c0 + i, k = k0 + threadIdx.x;
    if (cc < Cpad && k < Mp) {
      const float v = tile[threadIdx.x][i];
      const float t = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
      hi[(size_t)cc * Mp + k] = v;
      lo[(size_t)cc * Mp + k] = v - t;
    }
  }
}

__global__ void __launch_bounds__(kHtThreads, 1)
head_tail_umma_kernel(const __grid_constant__ HeadTailMaps tm, const float* __restrict__ bias,
                      float* const* __restrict__ dst, int row0, int R, int C, int K) {
  extern __shared__ __align__(1024) uint8_t ht_smem[];
  if ((ptx::smem_u32(ht_smem) & 1023u) != 0) __trap();
  uint64_t* full = reinterpret_cast<uint64_t*>(ht_smem + kHtStages * kHtStageBytes);
  uint64_t* empty = full + kHtStages;
  uint64_t* tmem_full = empty + kHtStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * kHtN, r0 = blockIdx.y * kHtM;
  const int ksteps = (K + kHtK - 1) / kHtK;
  pdl_trigger();
  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tensormap(&tm.a_hi); ptx::prefetch_tensormap(&tm.a_lo);
    ptx::prefetch_tensormap(&tm.b_hi); ptx::prefetch_tensormap(&tm.b_lo);
    for (int s = 0; s < kHtStages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    ptx::mbar_init(tmem_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<kHtN>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      // the weight planes do not depend on the head kernel: their first stages go out before the
      // dependency wait, the ê planes after it
      int ahead = 0;
      for (; ahead < kHtStages && ahead < ksteps; ++ahead) {
        uint8_t* st = ht_smem + ahead * kHtStageBytes;
        ptx::mbar_arrive_expect_tx(&full[ahead], kHtStageBytes);
        ptx::tma_load_2d(st + 2 * kHtTileBytes, &tm.b_hi, ahead * kHtK, c0, &full[ahead]);
        ptx::tma_load_2d(st + 3 * kHtTileBytes, &tm.b_lo, ahead * kHtK, c0, &full[ahead]);
      }
      pdl_wait();
      for (int ks = 0; ks < ksteps; ++ks) {
        uint8_t* st = ht_smem + stage * kHtStageBytes;
        if (ks >= ahead) {
          ptx::mbar_wait_bounded(&empty[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full[stage], kHtStageBytes);
          ptx::tma_load_2d(st + 2 * kHtTileBytes, &tm.b_hi, ks * kHtK, c0, &full[stage]);
          ptx::tma_load_2d(st + 3 * kHtTileBytes, &tm.b_lo, ks * kHtK, c0, &full[stage]);
        }
        ptx::tma_load_2d(st, &tm.a_hi, ks * kHtK, row0 + r0, &full[stage]);
        ptx::tma_load_2d(st + kHtTileBytes, &tm.a_lo, ks * kHtK, row0 + r0, &full[stage]);
        if (++stage == kHtStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_tf32(kHtM, kHtN);
      int stage = 0;
      uint32_t phase = 0;
      for (int ks = 0; ks < ksteps; ++ks) {
        ptx::mbar_wait_bounded(&full[stage], phase);
        ptx::tc_fence_after();
        const uint32_t base = ptx::smem_u32(ht_smem + stage * kHtStageBytes);
        const uint64_t a_hi = ptx::make_smem_desc_sw128(base);
        const uint64_t a_lo = ptx::make_smem_desc_sw128(base + kHtTileBytes);
        const uint64_t b_hi = ptx::make_smem_desc_sw128(base + 2 * kHtTileBytes);
        const uint64_t b_lo = ptx::make_smem_desc_sw128(base + 3 * kHtTileBytes);
#pragma unroll
        for (int k = 0; k < kHtK / 8; ++k) {   // 32 bytes (2 x 16-byte units) along K per step
          ptx::umma_tf32(tmem_d, a_lo + 2 * k, b_hi + 2 * k, idesc, (ks | k) != 0);
          ptx::umma_tf32(tmem_d, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
          ptx::umma_tf32(tmem_d, a_hi + 2 * k, b_hi + 2 * k, idesc, 1);
        }
        ptx::umma_commit(&empty[stage]);
        if (++stage == kHtStages) { stage = 0; phase ^= 1; }
      }
      ptx::umma_commit(tmem_full);
    }
  } else {
    const int quarter = warp & 3;
    const int r = r0 + quarter * 32 + lane;
    ptx::mbar_wait_bounded(tmem_full, 0);
    ptx::tc_fence_after();
    float* out = r < R ? dst[row0 + r] : nullptr;
    for (int cb = 0; cb < kHtN; cb += 32) {
      float v[32];
      ptx::tmem_ld_32x32b_x32(tmem_d + (static_cast<uint32_t>(quarter * 32) << 16) + cb, v);
      if (out != nullptr) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int cc = c0 + cb + c;
          if (cc < C) out[cc] = v[c] + bias[cc];
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kHtN>(tmem_d);
  }
}

}  // namespace n2nmn
