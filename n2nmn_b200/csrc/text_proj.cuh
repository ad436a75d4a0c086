// K1: text projection for every node that takes a text vector, and the quadratic-form
// coefficients of the Transform nodes.
//   tau[r, :] = word_vecs[t_r, b_r, :] · W_txt[set] + b_txt[set]
// (fc('fc_text') / fc('text_fc'): models_clevr/nmn3_modules.py:104,167,209,429,478 through
//  util/cnn.py:116). The gather of _slice_word_vecs (nmn3_modules.py:53-57) is folded into the
// load: row (t*N + b) of the segment's time-major word_vecs.
//
// Both are small dense products  C[r, c] = Σ_k A[r, k] · B[k, c]  with a few hundred to a few
// thousand rows per launch, done exactly in fp32 on the CUDA cores as one tiled kernel:
//   CTA tile 64 rows x 64 columns, K in chunks of 32 staged through shared memory (the next chunk
//   is in registers while the current one is multiplied), 256 threads = 16 column quads x 16 row
//   quads, 4 x 4 accumulators each: 2 shared-memory loads per 16 FMAs.
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
constexpr int kTileK = 32;       // K per shared-memory chunk
constexpr int kTileRows = kTextRowsPerCta;   // 64 rows per CTA
static_assert(kTextRowsPerCta == 64, "the tile kernel is written for 64-row groups");

// Rows of each text weight set, by value: a CTA finds its group without touching global memory.
struct TextSetRows { int32_t start[NUM_TEXT_SETS + 1]; };

// One 64 x 64 output tile. a_row(r) -> pointer to row r of A (or nullptr: zeros), K valid values
// per row; B row pitch ldb; sq_from_col: columns >= this use A² instead of A (quad kernel).
// acc[i][j] = C[4*ty + i][c0 + 4*tx + j].
template <class ARow>
__device__ __forceinline__ void tile_gemm_64x64(ARow a_row, int K, const float* __restrict__ B,
                                                int ldb, int c0, int ncols, int sq_from_col,
                                                float (&acc)[4][4]) {
  __shared__ __align__(16) float As[2][kTileK][kTileRows + 4];   // [k][row], padded
  __shared__ __align__(16) float Bs[2][kTileK][kTextCols];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  // loader roles: A: thread -> (row = tid / 4, k-quads (tid % 4) and +4); B: (k = tid / 8 ... )
  const int ar = threadIdx.x >> 2, aq = threadIdx.x & 3;          // 64 rows x 4 threads
  const int bk = threadIdx.x >> 4, bq = threadIdx.x & 15;         // 16 k-rows x 16 col quads
  const float* arow = a_row(ar);
  const bool a_vec = arow != nullptr && (reinterpret_cast<uintptr_t>(arow) & 15) == 0;
  const bool bcol_ok = c0 + 4 * bq < ncols;
  float4 ra[2], rb[2];
  auto load_chunk = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = k0 + 4 * (aq + 4 * h);
      ra[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (arow != nullptr) {
        if (k + 3 < K && a_vec) ra[h] = __ldg(reinterpret_cast<const float4*>(arow + k));
        else {
          if (k < K) ra[h].x = __ldg(arow + k);
          if (k + 1 < K) ra[h].y = __ldg(arow + k + 1);
          if (k + 2 < K) ra[h].z = __ldg(arow + k + 2);
          if (k + 3 < K) ra[h].w = __ldg(arow + k + 3);
        }
      }
      const int kb = k0 + bk + 16 * h;
      rb[h] = (kb < K && bcol_ok)
                  ? __ldg(reinterpret_cast<const float4*>(B + (size_t)kb * ldb + c0) + bq)
                  : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_chunk = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kk = 4 * (aq + 4 * h);
      As[buf][kk][ar] = ra[h].x; As[buf][kk + 1][ar] = ra[h].y;
      As[buf][kk + 2][ar] = ra[h].z; As[buf][kk + 3][ar] = ra[h].w;
      *reinterpret_cast<float4*>(&Bs[buf][bk + 16 * h][4 * bq]) = rb[h];
    }
  };
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool sq = c0 + 4 * tx >= sq_from_col;
  const int nchunks = (K + kTileK - 1) / kTileK;
  load_chunk(0);
  store_chunk(0);
  __syncthreads();
  for (int ch = 0; ch < nchunks; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < nchunks) load_chunk((ch + 1) * kTileK);   // in flight during the FMAs below
#pragma unroll
    for (int k = 0; k < kTileK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[buf][k][4 * ty]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][4 * tx]);
      if (sq) { a.x *= a.x; a.y *= a.y; a.z *= a.z; a.w *= a.w; }
      acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
      acc[0][2] = fmaf(a.x, b.z, acc[0][2]); acc[0][3] = fmaf(a.x, b.w, acc[0][3]);
      acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
      acc[1][2] = fmaf(a.y, b.z, acc[1][2]); acc[1][3] = fmaf(a.y, b.w, acc[1][3]);
      acc[2][0] = fmaf(a.z, b.x, acc[2][0]); acc[2][1] = fmaf(a.z, b.y, acc[2][1]);
      acc[2][2] = fmaf(a.z, b.z, acc[2][2]); acc[2][3] = fmaf(a.z, b.w, acc[2][3]);
      acc[3][0] = fmaf(a.w, b.x, acc[3][0]); acc[3][1] = fmaf(a.w, b.y, acc[3][1]);
      acc[3][2] = fmaf(a.w, b.z, acc[3][2]); acc[3][3] = fmaf(a.w, b.w, acc[3][3]);
    }
    if (ch + 1 < nchunks) {
      store_chunk(buf ^ 1);     // the other buffer: its readers finished before the last barrier
      __syncthreads();
    }
  }
}

// grid = (Mp / 64, groups); one group = <= 64 text rows of ONE weight set (schedule.cpp)
__global__ void __launch_bounds__(256)
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
  float acc[4][4];
  tile_gemm_64x64(a_row, md.Dt, md.txt_w[g.set], Mp, c0, Mp, 1 << 30, acc);
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
  for (int i = 0; i < 4; ++i) {
    const int r = 4 * ty + i;
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
__global__ void __launch_bounds__(256)
quad_kernel(DevModel md, TextBufs tb, int row0, int nrows) {
  pdl_trigger();
  pdl_wait();      // tau comes from the text kernel
  const int Mp = md.Mp, qp = quad_pitch(md.ksize);
  const int r0 = row0 + blockIdx.y * kTileRows, cnt = min(kTileRows, row0 + nrows - r0);
  const int c0 = blockIdx.x * kTextCols;
  auto a_row = [&](int r) -> const float* {
    return r < cnt ? tb.tau + (size_t)(r0 + r) * Mp : nullptr;
  };
  float acc[4][4];
  // columns [0, n) are u (pairs with tau), padded to a multiple of 4; the rest is Q (tau²)
  tile_gemm_64x64(a_row, Mp, md.conv_quad, qp, c0, qp, quad_u_pitch(md.ksize), acc);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int c = c0 + 4 * tx;
  if (c >= qp) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 4 * ty + i;
    if (r < cnt)
      *reinterpret_cast<float4*>(tb.tq + (size_t)(r0 + r) * qp + c) =
          make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
}

}  // namespace n2nmn
