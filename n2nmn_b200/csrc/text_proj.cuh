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
#include "tile_gemm.cuh"

namespace n2nmn {

// Rows of each text weight set, by value: a CTA finds its group without touching global memory.
struct TextSetRows { int32_t start[NUM_TEXT_SETS + 1]; };

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
