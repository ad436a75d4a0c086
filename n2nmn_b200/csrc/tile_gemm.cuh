// 64 x 64 fp32 tile GEMM with the whole K extent staged in shared memory by cp.async: the inner
// product engine of the text projection / quadratic-form kernels (text_proj.cuh) and of the
// seq2seq LSTM steps (seq2seq.cu). See text_proj.cuh for the design notes.
#pragma once
#include "common.cuh"

namespace n2nmn {

constexpr int kTextCols = 64;    // output columns per CTA
constexpr int kTileRows = kTextRowsPerCta;   // 64 rows per CTA
constexpr int kTileThreads = 512;
static_assert(kTextRowsPerCta == 64, "the tile kernel is written for 64-row groups");

// cp.async helpers (16-byte copies global -> shared, completion by commit groups)
__device__ __forceinline__ void tp_cp16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
               ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void tp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kTileKBlock = 320;   // K extent staged at a time (Dt = 300 and Mp = 256 in one pass)
__host__ __device__ inline int tile_a_pitch(int K) {   // floats; 16-byte rows
  return ((K < kTileKBlock ? K : kTileKBlock) + 3) & ~3;
}
__host__ __device__ inline int tile_smem_floats(int K) {
  return kTileRows * tile_a_pitch(K) + tile_a_pitch(K) * kTextCols;
}

// One 64 x 64 output tile. The whole A tile [64][K] and B tile [K][64] are brought into shared
// memory with cp.async in TWO commit groups (first / second half of K): every load of the tile is
// in flight at once (the launch has at most a CTA or two per SM, so nothing else hides the
// latency), and the FMAs of the first half run under the second half's loads.
// a_row(r) -> pointer to row r of A (or nullptr: zeros), K valid values per row; B row pitch ldb;
// sq_from_col: columns >= this use A² instead of A (quad kernel).
// acc[i][j] = C[2*ty + i][c0 + 4*tx + j]. smem: tile_smem_floats(K) floats + 64 pointers.
template <class ARow>
__device__ __forceinline__ void tile_gemm_64x64(float* smem, ARow a_row, int Ktot,
                                                const float* __restrict__ Btot, int ldb, int c0,
                                                int ncols, int sq_from_col, float (&acc)[2][4],
                                                bool zero_acc = true) {
  const int P = tile_a_pitch(Ktot);
  float* As = smem;                      // [64][P]
  float* Bs = smem + kTileRows * P;      // [P][64]
  const float** s_ap = reinterpret_cast<const float**>(Bs + P * kTextCols);   // [64]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  if (threadIdx.x < kTileRows) s_ap[threadIdx.x] = a_row(threadIdx.x);
  if (zero_acc) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  }
  const bool sq = c0 + 4 * tx >= sq_from_col;
  for (int kb0 = 0; kb0 < Ktot; kb0 += kTileKBlock) {   // (one pass unless K > kTileKBlock)
  const int K = min(kTileKBlock, Ktot - kb0), K4 = (K + 3) & ~3;
  const float* __restrict__ B = Btot + (size_t)kb0 * ldb;
  __syncthreads();                        // s_ap visible / the previous pass is done with smem
  const int Kh = ((K4 / 2) + 3) & ~3;    // first half: k in [0, Kh)
  const int qa = P >> 2;                 // 16-byte quads per A row
  for (int half = 0; half < 2; ++half) {
    const int k_lo = half ? Kh : 0, k_hi = half ? K4 : Kh;
    for (int i = threadIdx.x; i < kTileRows * qa; i += blockDim.x) {
      const int r = i / qa, k = 4 * (i - r * qa);
      if (k < k_lo || k >= k_hi) continue;
      const float* src = s_ap[r] ? s_ap[r] + kb0 : nullptr;
      float* dst = As + r * P + k;
      if (src != nullptr && k + 3 < K && ((reinterpret_cast<uintptr_t>(src + k) & 15) == 0)) {
        tp_cp16(dst, src + k);
      } else {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src != nullptr) {
          if (k < K) v.x = __ldg(src + k);
          if (k + 1 < K) v.y = __ldg(src + k + 1);
          if (k + 2 < K) v.z = __ldg(src + k + 2);
          if (k + 3 < K) v.w = __ldg(src + k + 3);
        }
        *reinterpret_cast<float4*>(dst) = v;
      }
    }
    for (int i = threadIdx.x; i < (k_hi - k_lo) * (kTextCols / 4); i += blockDim.x) {
      const int k = k_lo + i / (kTextCols / 4), q = i % (kTextCols / 4);
      float* dst = Bs + k * kTextCols + 4 * q;
      if (k < K && c0 + 4 * q < ncols) tp_cp16(dst, B + (size_t)k * ldb + c0 + 4 * q);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    tp_commit();
  }
  for (int half = 0; half < 2; ++half) {
    if (half == 0) tp_wait<1>(); else tp_wait<0>();
    __syncthreads();
    const int k_lo = half ? Kh : 0, k_hi = half ? K4 : Kh;
#pragma unroll 2
    for (int k = k_lo; k < k_hi; k += 4) {
      float4 a[2], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        b[i] = *reinterpret_cast<const float4*>(Bs + (k + i) * kTextCols + 4 * tx);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        a[i] = *reinterpret_cast<const float4*>(As + (2 * ty + i) * P + k);
        if (sq) { a[i].x *= a[i].x; a[i].y *= a[i].y; a[i].z *= a[i].z; a[i].w *= a[i].w; }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float ak[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          acc[i][0] = fmaf(ak[kk], b[kk].x, acc[i][0]); acc[i][1] = fmaf(ak[kk], b[kk].y, acc[i][1]);
          acc[i][2] = fmaf(ak[kk], b[kk].z, acc[i][2]); acc[i][3] = fmaf(ak[kk], b[kk].w, acc[i][3]);
        }
      }
    }
  }
  }
}

}  // namespace n2nmn
