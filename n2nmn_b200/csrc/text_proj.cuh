// K1: text projection for every node that takes a text vector, and the quadratic-form
// coefficients of the Transform nodes.
//   tau[r, :] = word_vecs[t_r, b_r, :] · W_txt[set] + b_txt[set]
// (fc('fc_text') / fc('text_fc'): models_clevr/nmn3_modules.py:104,167,209,429,478 through
//  util/cnn.py:116). The gather of _slice_word_vecs (nmn3_modules.py:53-57) is folded into the
// load: row (t*N + b) of the segment's time-major word_vecs.
//
// Both are small dense products  C[r, c] = Σ_k A[r, k] · B[k, c]  with a few hundred to a few
// thousand rows per launch, done exactly in fp32 on the CUDA cores as one tiled kernel:
//   CTA tile 64 rows x 64 columns with the whole K extent of both operands staged in shared
//   memory by cp.async, 512 threads = 16 column quads x 32 row pairs, 2 x 4 accumulators each
//   (a launch has at most one CTA per SM, so the warps that hide the shared-memory latency have
//   to come from inside the CTA: ncu r2f showed 12 % warp occupancy and 42 K cycles per tile with
//   256 threads x 4 x 4).
// Round 1 gave every group of 8 rows its own CTA, which streamed the whole weight matrix from L2
// with a handful of loads in flight: ~1000 SM-cycles per row against ~600 of FMA issue at peak.
//   text_proj_kernel : A = gathered word vectors (K = Dt), B = W_txt [Dt][Mp]; emits tau,
//                      tau∘w_eltwise and tau² so the consumers' epilogues are pure FMAs.
//   quad_kernel      : A = tau (first n outputs) or tau² (the rest) of the Transform rows, K = Mp,
//                      B = conv_quad^T [Mp][quad_pitch] (common.cuh); emits (u, Q) per node.
// Weights are stored with row pitch Mp (zero padded), so padded columns come out as exact zeros.
#pragma once
#include "common.cuh"

namespace n2nmn {

constexpr int kTextCols = 64;    // output columns per CTA
constexpr int kTileRows = kTextRowsPerCta;   // 64 rows per CTA
constexpr int kTileThreads = 512;
static_assert(kTextRowsPerCta == 64, "the tile kernel is written for 64-row groups");

// Rows of each text weight set, by value: a CTA finds its group without touching global memory.
struct TextSetRows { int32_t start[NUM_TEXT_SETS + 1]; };

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
                                                int ncols, int sq_from_col, float (&acc)[2][4]) {
  const int P = tile_a_pitch(Ktot);
  float* As = smem;                      // [64][P]
  float* Bs = smem + kTileRows * P;      // [P][64]
  const float** s_ap = reinterpret_cast<const float**>(Bs + P * kTextCols);   // [64]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  if (threadIdx.x < kTileRows) s_ap[threadIdx.x] = a_row(threadIdx.x);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
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

// grid = (Mp / 64, groups); one group = <= 64 text rows of ONE weight set (schedule.cpp)
__global__ void __launch_bounds__(kTileThreads)
text_proj_kernel(DevModel md, TextBufs tb, TextSetRows rows,
                 const int32_t* __restrict__ text_t, const int32_t* __restrict__ text_b) {
  pdl_trigger();   // the contraction kernel only needs our output in its epilogue
  if (threadIdx.x == 0) N2NMN_STAMP(0, 0);
  const int M = md.M, Mp = md.Mp;
  TextGroup g;   // blockIdx.y-th group, groups never straddle weight sets
  {
    int gi = blockIdx.y, set = 0;
    for (; set < NUM_TEXT_SETS; ++set) {
      const int ng = (rows.start[set + 1] - rows.start[set] + kTextRowsPerCta - 1) / kTextRowsPerCta;
      if (gi < ng) break;
      gi -= ng;
    }
    g.set = set;
    g.start = rows.start[set] + gi * kTextRowsPerCta;
    g.count = min(kTextRowsPerCta, rows.start[set + 1] - g.start);
  }
  const int es = (g.set == TS_FIND) ? ES_FIND : (g.set == TS_FSP) ? ES_FSP
               : (g.set == TS_TRANSFORM) ? ES_TRANSFORM : -1;
  const int c0 = blockIdx.x * kTextCols;
  auto a_row = [&](int r) -> const float* {
    return r < g.count ? word_vec_row(md, text_t[g.start + r], text_b[g.start + r]) : nullptr;
  };
  extern __shared__ __align__(16) float tile_smem[];
  float acc[2][4];
  tile_gemm_64x64(tile_smem, a_row, md.Dt, md.txt_w[g.set], Mp, c0, Mp, 1 << 30, acc);
  if (threadIdx.x == 0) N2NMN_STAMP(0, 4);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int c = c0 + 4 * tx;
  float bias[4], w2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool live = c + j < M;
    bias[j] = live ? md.txt_b[g.set][c + j] : 0.f;
    w2[j] = (live && es >= 0) ? md.elt_w[es][c + j] : 1.f;
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = 2 * ty + i;
    if (r >= g.count) continue;
    float4 v, vw, v2;
    float* pv = &v.x; float* pw = &vw.x; float* p2 = &v2.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float t = (c + j < M) ? acc[i][j] + bias[j] : 0.f;
      pv[j] = t; pw[j] = t * w2[j]; p2[j] = t * t;
    }
    const size_t idx = (size_t)(g.start + r) * Mp + c;
    *reinterpret_cast<float4*>(tb.tau + idx) = v;
    *reinterpret_cast<float4*>(tb.tauw + idx) = vw;
    *reinterpret_cast<float4*>(tb.tau2 + idx) = v2;
  }
  if (threadIdx.x == 0) N2NMN_STAMP(0, 5);
}

// (u, Q) of the Transform nodes (common.cuh): tq[row, o] = Σ_c (o < n ? tau : tau²)[row, c] ·
// conv_quad^T[c, o] for the text rows [row0, row0 + nrows) of the Transform weight set.
// grid = (ceil(quad_pitch / 64), ceil(nrows / 64)).
__global__ void __launch_bounds__(kTileThreads)
quad_kernel(DevModel md, TextBufs tb, int row0, int nrows) {
  pdl_trigger();
  pdl_wait();      // tau comes from the text kernel
  const int Mp = md.Mp, qp = quad_pitch(md.ksize);
  const int r0 = row0 + blockIdx.y * kTileRows, cnt = min(kTileRows, row0 + nrows - r0);
  const int c0 = blockIdx.x * kTextCols;
  auto a_row = [&](int r) -> const float* {
    return r < cnt ? tb.tau + (size_t)(r0 + r) * Mp : nullptr;
  };
  extern __shared__ __align__(16) float tile_smem[];
  float acc[2][4];
  // columns [0, n) are u (pairs with tau), padded to a multiple of 4; the rest is Q (tau²)
  tile_gemm_64x64(tile_smem, a_row, Mp, md.conv_quad, qp, c0, qp, quad_u_pitch(md.ksize), acc);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int c = c0 + 4 * tx;
  if (c >= qp) return;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = 2 * ty + i;
    if (r < cnt)
      *reinterpret_cast<float4*>(tb.tq + (size_t)(r0 + r) * qp + c) =
          make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
}

}  // namespace n2nmn
