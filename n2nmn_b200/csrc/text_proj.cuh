// K1: text projection for every node that takes a text vector.
//   tau[r, :] = word_vecs[t_r, b_r, :] · W_txt[set] + b_txt[set]
// (fc('fc_text') / fc('text_fc'): models_clevr/nmn3_modules.py:104,167,209,429,478 through
//  util/cnn.py:116). The gather of _slice_word_vecs (nmn3_modules.py:53-57) is folded into the
// load: row (t*N + b) of the time-major word_vecs.
//
// One CTA = up to 8 nodes of ONE weight set x the 64-column blocks blockIdx.x, blockIdx.x +
// gridDim.x, ... (gridDim.x = Mp/64: one block per CTA, shortest kernel; gridDim.x = 1: one CTA
// per node group walks all column blocks and gathers its word vectors once — fewer, longer CTAs,
// less SM-time, used when many batches are in flight). The 256 threads form a
// 16 (column quads) x 16 (K slices) grid: every thread streams ~Dt/16 float4 weight rows with all
// loads independent (the kernel is latency-bound, so memory-level parallelism is what matters),
// keeps 8x4 accumulators, and the 16 K slices are reduced through shared memory.
// Also emits tau∘w_eltwise and tau² so the consumers' epilogues are pure FMAs, and — for the
// rows of TransformModule — the node's quadratic-form coefficients (common.cuh) as a second small
// product [8 rows x M] x [M x quad_rows] against md.conv_quad; a Transform group is therefore
// always handled by ONE CTA (it needs the complete tau rows).
// Weights are stored with row pitch Mp (zero padded), so padded columns come out as exact zeros.
#pragma once
#include "common.cuh"

namespace n2nmn {

constexpr int kTextCols = 64;   // output columns per CTA

constexpr int kTextKIter = 20;   // K rows per thread per chunk (16 slices x 20 = 320 >= Dt=300)

// Rows of each text weight set, by value: a CTA finds its group without touching global memory,
// so its weight loads can start immediately.
struct TextSetRows { int32_t start[NUM_TEXT_SETS + 1]; };

__global__ void __launch_bounds__(256)
text_proj_kernel(DevModel md, TextBufs tb, TextSetRows rows,
                 const int32_t* __restrict__ text_t, const int32_t* __restrict__ text_b) {
  extern __shared__ float s_dyn[];
  __shared__ const float* s_src[kTextRowsPerCta];
  pdl_trigger();   // the contraction kernel only needs our output in its epilogue
  if (threadIdx.x == 0) N2NMN_STAMP(0, 0);
  const int Dt = md.Dt, M = md.M, Mp = md.Mp;
  float* s_x = s_dyn;                                 // [8][Dt]
  float* s_red = s_dyn + kTextRowsPerCta * Dt;        // [8 warps][8 rows][64 cols]
  float* s_tau = s_red + 8 * kTextRowsPerCta * kTextCols;   // [8][Mp] tau, then [8][Mp] tau²
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tx = lane & 15;                    // column quad inside the CTA's 64 columns
  const int ky = warp * 2 + (lane >> 4);       // K slice 0..15
  const int n_cblk = Mp / kTextCols;
  int cblk = blockIdx.x;
  int c0 = cblk * kTextCols + tx * 4;
  TextGroup g;   // blockIdx.y-th group of <= 8 rows, groups never straddle weight sets
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
  // a Transform group: one CTA walks every column block (it needs whole tau rows afterwards)
  const bool quad = (g.set == TS_TRANSFORM) && md.conv_quad != nullptr;
  if (quad && blockIdx.x != 0) return;
  const int cb_step = quad ? 1 : gridDim.x;
  const float* __restrict__ wbase = md.txt_w[g.set] + c0;
  const int es = (g.set == TS_FIND) ? ES_FIND : (g.set == TS_FSP) ? ES_FSP
               : (g.set == TS_TRANSFORM) ? ES_TRANSFORM : -1;

  // (1) this thread's weight rows of the first chunk: independent of everything else, so the
  //     loads fly while the word vectors are being gathered
  float4 w[kTextKIter];
#pragma unroll
  for (int j = 0; j < kTextKIter; ++j) {
    const int k = ky + 16 * j;
    w[j] = (k < Dt) ? __ldg(reinterpret_cast<const float4*>(wbase + (size_t)k * Mp))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (threadIdx.x == 0) N2NMN_STAMP(0, 1);
  // (2) source rows of the group's nodes in the time-major word_vecs: row t*N + b of the segment
  if (threadIdx.x < kTextRowsPerCta) {
    const int r = threadIdx.x;
    s_src[r] = (r < g.count) ? word_vec_row(md, text_t[g.start + r], text_b[g.start + r]) : nullptr;
  }
  __syncthreads();
  if (threadIdx.x == 0) N2NMN_STAMP(0, 2);
  // (3) gather the word vectors (all loads of a thread are independent)
  constexpr int kGather = 10;   // 8 rows x 300 words / 256 threads = 9.4 loads per thread
  for (int i0 = 0; i0 < kTextRowsPerCta * Dt; i0 += kGather * 256) {
    float xv[kGather];
#pragma unroll
    for (int u = 0; u < kGather; ++u) {
      const int i = i0 + u * 256 + threadIdx.x;
      xv[u] = 0.f;
      if (i < kTextRowsPerCta * Dt) {
        const int r = i / Dt, k = i - r * Dt;
        const float* src = s_src[r];
        if (src != nullptr) xv[u] = __ldg(src + k);
      }
    }
#pragma unroll
    for (int u = 0; u < kGather; ++u) {
      const int i = i0 + u * 256 + threadIdx.x;
      if (i < kTextRowsPerCta * Dt) s_x[i] = xv[u];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) N2NMN_STAMP(0, 3);
  for (; cblk < n_cblk; cblk += cb_step) {
  c0 = cblk * kTextCols + tx * 4;
  const float* __restrict__ wb = md.txt_w[g.set] + c0;
  // (4) 8 x 4 accumulators per thread
  float4 acc[kTextRowsPerCta];
#pragma unroll
  for (int r = 0; r < kTextRowsPerCta; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int kb = 0; kb < Dt; kb += 16 * kTextKIter) {
    if (kb > 0 || cblk != (int)blockIdx.x) {   // chunk / column block not prefetched above
#pragma unroll
      for (int j = 0; j < kTextKIter; ++j) {
        const int k = kb + ky + 16 * j;
        w[j] = (k < Dt) ? __ldg(reinterpret_cast<const float4*>(wb + (size_t)k * Mp))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int j = 0; j < kTextKIter; ++j) {
      const int k = kb + ky + 16 * j;
      if (k < Dt) {
#pragma unroll
        for (int r = 0; r < kTextRowsPerCta; ++r) {
          const float x = s_x[r * Dt + k];
          acc[r].x = fmaf(x, w[j].x, acc[r].x); acc[r].y = fmaf(x, w[j].y, acc[r].y);
          acc[r].z = fmaf(x, w[j].z, acc[r].z); acc[r].w = fmaf(x, w[j].w, acc[r].w);
        }
      }
    }
  }
  if (threadIdx.x == 0) N2NMN_STAMP(0, 4);
  // the two K slices inside a warp, then the 8 warps through shared memory
#pragma unroll
  for (int r = 0; r < kTextRowsPerCta; ++r) {
    acc[r].x += __shfl_xor_sync(0xffffffffu, acc[r].x, 16);
    acc[r].y += __shfl_xor_sync(0xffffffffu, acc[r].y, 16);
    acc[r].z += __shfl_xor_sync(0xffffffffu, acc[r].z, 16);
    acc[r].w += __shfl_xor_sync(0xffffffffu, acc[r].w, 16);
    if (lane < 16)
      *reinterpret_cast<float4*>(s_red + ((warp * kTextRowsPerCta + r) * kTextCols + tx * 4)) =
          acc[r];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < kTextRowsPerCta * kTextCols; o += blockDim.x) {
    const int r = o / kTextCols, cc = o - r * kTextCols;
    if (r >= g.count) continue;
    const int c = cblk * kTextCols + cc;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += s_red[(w * kTextRowsPerCta + r) * kTextCols + cc];
    const bool live = c < M;
    v = live ? v + md.txt_b[g.set][c] : 0.f;
    const float w2 = (live && es >= 0) ? md.elt_w[es][c] : 1.f;
    const size_t idx = (size_t)(g.start + r) * Mp + c;
    tb.tau[idx] = v;
    tb.tauw[idx] = v * w2;
    tb.tau2[idx] = v * v;
    if (quad) { s_tau[r * Mp + c] = v; s_tau[(kTextRowsPerCta + r) * Mp + c] = v * v; }
  }
  __syncthreads();   // s_red is reused by the next column block
  }
  if (quad) {
    // (u, Q) of the group's Transform nodes: out[r][o] = Σ_c conv_quad[o][c] · (o < n ? tau : tau²)[r][c]
    const int nq = quad_rows(md.ksize), n1 = quad_n(md.ksize), qp = quad_pitch(md.ksize);
    for (int o = threadIdx.x; o < nq; o += blockDim.x) {
      const float4* __restrict__ krow = reinterpret_cast<const float4*>(md.conv_quad + (size_t)o * Mp);
      const float* tv = s_tau + (o < n1 ? 0 : kTextRowsPerCta * Mp);
      float acc[kTextRowsPerCta];
#pragma unroll
      for (int r = 0; r < kTextRowsPerCta; ++r) acc[r] = 0.f;
#pragma unroll 4
      for (int c4 = 0; c4 < (Mp >> 2); ++c4) {
        const float4 k4 = __ldg(krow + c4);
#pragma unroll
        for (int r = 0; r < kTextRowsPerCta; ++r) {
          const float4 t4 = *reinterpret_cast<const float4*>(tv + r * Mp + 4 * c4);
          acc[r] = fmaf(k4.x, t4.x, acc[r]); acc[r] = fmaf(k4.y, t4.y, acc[r]);
          acc[r] = fmaf(k4.z, t4.z, acc[r]); acc[r] = fmaf(k4.w, t4.w, acc[r]);
        }
      }
      for (int r = 0; r < g.count; ++r) tb.tq[(size_t)(g.start + r) * qp + o] = acc[r];
    }
  }
  if (threadIdx.x == 0) N2NMN_STAMP(0, 5);
}

}  // namespace n2nmn
