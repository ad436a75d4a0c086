// K1: text projection for every node that takes a text vector.
//   tau[r, :] = word_vecs[t_r, b_r, :] · W_txt[set] + b_txt[set]
// (fc('fc_text') / fc('text_fc'): models_clevr/nmn3_modules.py:104,167,209,429,478 through
//  util/cnn.py:116). The gather of _slice_word_vecs (nmn3_modules.py:53-57) is folded into the
// load: row (t*N + b) of the time-major word_vecs.
//
// One CTA = up to 8 nodes of ONE weight set x 256 output columns; the weight matrix row is read
// once per CTA (coalesced over columns) and reused for the 8 nodes from registers.
// Also emits tau∘w_eltwise and tau² so the consumers' epilogues are pure FMAs.
#pragma once
#include "common.cuh"

namespace n2nmn {

__global__ void __launch_bounds__(256)
text_proj_kernel(DevModel md, TextBufs tb, const TextGroup* __restrict__ groups,
                 const int32_t* __restrict__ text_t, const int32_t* __restrict__ text_b) {
  extern __shared__ float s_x[];  // [8][Dt]
  const TextGroup g = groups[blockIdx.y];
  const int Dt = md.Dt, M = md.M, Mp = md.Mp;
  for (int i = threadIdx.x; i < g.count * Dt; i += blockDim.x) {
    const int r = i / Dt, k = i - r * Dt;
    const int row = g.start + r;
    s_x[r * Dt + k] = md.word_vecs[((size_t)text_t[row] * md.N + text_b[row]) * Dt + k];
  }
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Mp) return;
  float acc[kTextRowsPerCta];
  const bool live = c < M;
  const float bias = live ? md.txt_b[g.set][c] : 0.f;
#pragma unroll
  for (int r = 0; r < kTextRowsPerCta; ++r) acc[r] = bias;
  if (live) {
    const float* __restrict__ w = md.txt_w[g.set] + c;
#pragma unroll 4
    for (int k = 0; k < Dt; ++k) {
      const float wk = __ldg(w + (size_t)k * M);
#pragma unroll
      for (int r = 0; r < kTextRowsPerCta; ++r) acc[r] = fmaf(s_x[r * Dt + k], wk, acc[r]);
    }
  }
  const int es = (g.set == TS_FIND) ? ES_FIND : (g.set == TS_FSP) ? ES_FSP
               : (g.set == TS_TRANSFORM) ? ES_TRANSFORM : -1;
  const float w2 = (live && es >= 0) ? md.elt_w[es][c] : 1.f;
#pragma unroll
  for (int r = 0; r < kTextRowsPerCta; ++r) {
    if (r < g.count) {
      const size_t o = (size_t)(g.start + r) * Mp + c;
      const float v = live ? acc[r] : 0.f;
      tb.tau[o] = v;
      tb.tauw[o] = v * w2;
      tb.tau2[o] = v * v;
    }
  }
}

}  // namespace n2nmn
