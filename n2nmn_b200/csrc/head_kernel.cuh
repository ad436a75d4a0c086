// K4: batched answer heads of the attention-pooled root modules (evaluation schedules with
// pooled_direct, schedule.hpp).
//
// DescribeModule / SamePropertyModule (models_clevr/nmn3_modules.py:402-495, VQA
// models_vqa/nmn3_modules.py:193-240) are
//     f   = Σ_p softmax(att)_p · X_b[p,:]                      (pooled feature, D values)
//     φ   = f · W_att + b_att            (fc_att; SameProperty: φ0, φ1 with their own W_att)
//     e   = τ ∘ φ (∘ φ1),  ê = l2_normalize(e),  scores = ê · W_out + b_out.
// Round 1 computed φ from a stored per-image map G = X_b·W_att + b made by the contraction kernel
// (Σ_p s_p = 1): that made HALF of the contraction kernel's tiles serve these few root nodes
// (two 128-row tiles of 16.8 MFMA each per image and layer, against 0.2 MFMA for f and φ done
// directly). Now the tree kernel writes f (one 150-row weighted sum of feature rows per node) and
// this kernel does the two small dense products for MANY root nodes at once, so every weight
// matrix is read once per <= 16 nodes instead of once per node (VQA: W_out is 12 MB).
//
// One CTA = up to kHeadNodesMax root nodes of one type (HeadWork). Exact fp32 on the CUDA cores:
// 100 nodes x 512 x 250 MACs per launch is microseconds of work, and fp32 here is tighter than the
// TF32 maps it replaces.
#pragma once
#include "node_eval.cuh"

namespace n2nmn {

constexpr int kHeadThreads = 256;

struct HeadSmem { int f, phi, scratch, total; };
__host__ __device__ inline HeadSmem head_smem_layout(int nn, int pitch, int Mp) {
  HeadSmem s;
  s.f = nn * pitch;                  // pooled rows of the chunk (nn = nodes per CTA)
  s.phi = 2 * nn * Mp;               // φ0 / φ1 (ê is written over φ0)
  s.scratch = 8 * kHeadNodesMax * 32;
  s.total = s.f + s.phi + s.scratch;
  return s;
}

// phi[n][c] = bias[c] + Σ_k F[n][k]·W[k*M + c] for the chunk's rows (rows >= cnt are zeros).
template <int NN>
__device__ __forceinline__ void head_fc_att(const float* __restrict__ Fs, int pitch, int Dk,
                                            const float* __restrict__ W, int M, int Mp,
                                            const float* __restrict__ bias, float* phi) {
  for (int c = threadIdx.x; c < Mp; c += kHeadThreads) {
    float acc[NN];
#pragma unroll
    for (int n = 0; n < NN; ++n) acc[n] = 0.f;
    if (c < M) {
      const float* __restrict__ w = W + c;
      int k = 0;
      for (; k + 4 <= Dk; k += 4) {
        const float w0 = __ldg(w + (size_t)k * M), w1 = __ldg(w + (size_t)(k + 1) * M),
                    w2 = __ldg(w + (size_t)(k + 2) * M), w3 = __ldg(w + (size_t)(k + 3) * M);
#pragma unroll
        for (int n = 0; n < NN; ++n) {
          const float4 f = *reinterpret_cast<const float4*>(Fs + n * pitch + k);   // broadcast
          acc[n] = fmaf(f.x, w0, acc[n]); acc[n] = fmaf(f.y, w1, acc[n]);
          acc[n] = fmaf(f.z, w2, acc[n]); acc[n] = fmaf(f.w, w3, acc[n]);
        }
      }
      for (; k < Dk; ++k) {
        const float w0 = __ldg(w + (size_t)k * M);
#pragma unroll
        for (int n = 0; n < NN; ++n) acc[n] = fmaf(Fs[n * pitch + k], w0, acc[n]);
      }
      const float b = bias[c];
#pragma unroll
      for (int n = 0; n < NN; ++n) acc[n] += b;
    }
#pragma unroll
    for (int n = 0; n < NN; ++n) phi[n * Mp + c] = acc[n];   // zero beyond M
  }
}

template <int NN>
__global__ void __launch_bounds__(kHeadThreads)
head_kernel(const NodeCtx c, const NodeRec* __restrict__ nodes,
            const HeadWork* __restrict__ work, const int32_t* __restrict__ list) {
  extern __shared__ __align__(16) float head_smem[];
  const DevModel& md = c.md;
  const int M = md.M, Mp = md.Mp, C = md.C, Dk = md.Dk;
  const float* __restrict__ pooled = c.pooled;
  const int pool_pitch = c.pool_pitch;
  const HeadSmem L = head_smem_layout(NN, pool_pitch, Mp);
  float* Fs = head_smem;
  float* phi0 = Fs + L.f;
  float* phi1 = phi0 + NN * Mp;
  float* scratch = phi1 + NN * Mp;
  __shared__ int s_node[kHeadNodesMax];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  const HeadWork wk = work[blockIdx.x];   // launch tables: uploaded before any kernel of the step
  const int cnt = wk.count;
  const bool two = (wk.op == OP_SAME_PROPERTY);
  if (threadIdx.x < kHeadNodesMax)
    s_node[threadIdx.x] = threadIdx.x < cnt ? list[wk.first + threadIdx.x] : -1;
  pdl_wait();            // the pooled rows come from the tree kernel, tau from the text kernel
  __syncthreads();

  for (int which = 0; which < (two ? 2 : 1); ++which) {
    // ---- the chunk's pooled rows -> shared memory (unused rows = 0)
    for (int i = threadIdx.x; i < NN * (pool_pitch >> 2); i += kHeadThreads) {
      const int n = i / (pool_pitch >> 2), q = i - n * (pool_pitch >> 2);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < cnt) {
        const NodeRec& nd = nodes[s_node[n]];
        const int row = which ? nd.aux2 : nd.aux;
        v = __ldg(reinterpret_cast<const float4*>(pooled + (size_t)row * pool_pitch) + q);
      }
      reinterpret_cast<float4*>(Fs + n * pool_pitch)[q] = v;
    }
    __syncthreads();
    const int set = two ? (which ? PS_SP_ATT1 : PS_SP_ATT0) : PS_DESC_ATT;
    head_fc_att<NN>(Fs, pool_pitch, Dk, md.proj_w[set], M, Mp, md.proj_b[set],
                    which ? phi1 : phi0);
    __syncthreads();
  }

  // ---- e = τ∘φ0(∘φ1), l2_normalize over the M channels (nmn3_modules.py:448, 491): a warp per node
  for (int n = warp; n < cnt; n += kHeadThreads / 32) {
    const NodeRec& nd = nodes[s_node[n]];
    const float* tau = c.tb.tau + (size_t)nd.text * Mp;
    float ss = 0.f;
    for (int ch = lane; ch < Mp; ch += 32) {
      float e = 0.f;
      if (ch < M) {
        e = tau[ch] * phi0[n * Mp + ch];
        if (two) e *= phi1[n * Mp + ch];
      }
      phi0[n * Mp + ch] = e;
      ss = fmaf(e, e, ss);
    }
    ss = warp_sum(ss);
    const float inv = rsqrtf(fmaxf(ss, kEps));
    for (int ch = lane; ch < Mp; ch += 32) phi0[n * Mp + ch] *= inv;
  }
  __syncthreads();

  // ---- scores = ê·W_out + b_out (fc_eltwise)
  const int os = two ? OS_SAMEPROP : OS_DESCRIBE;
  const float* __restrict__ Wo = md.out_w[os];
  const float* __restrict__ bo = md.out_b[os];
  if (C <= 32) {
    // warp w takes the channels k ≡ w (mod 8), lane = class; partial sums meet in `scratch`
    float acc[NN];
#pragma unroll
    for (int n = 0; n < NN; ++n) acc[n] = 0.f;
    if (lane < C) {
      for (int k = warp; k < M; k += kHeadThreads / 32) {
        const float w = __ldg(Wo + (size_t)k * C + lane);
#pragma unroll
        for (int n = 0; n < NN; ++n) acc[n] = fmaf(phi0[n * Mp + k], w, acc[n]);
      }
    }
#pragma unroll
    for (int n = 0; n < NN; ++n) scratch[(warp * kHeadNodesMax + n) * 32 + lane] = acc[n];
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * 32; i += kHeadThreads) {
      const int n = i >> 5, cl = i & 31;
      if (cl < C) {
        float v = bo[cl];
#pragma unroll
        for (int w = 0; w < kHeadThreads / 32; ++w) v += scratch[(w * kHeadNodesMax + n) * 32 + cl];
        score_row(c, nodes[s_node[n]].out)[cl] = v;
      }
    }
  } else {
    for (int cl = threadIdx.x; cl < C; cl += kHeadThreads) {
      float acc[NN];
#pragma unroll
      for (int n = 0; n < NN; ++n) acc[n] = 0.f;
#pragma unroll 4
      for (int k = 0; k < M; ++k) {
        const float w = __ldg(Wo + (size_t)k * C + cl);
#pragma unroll
        for (int n = 0; n < NN; ++n) acc[n] = fmaf(phi0[n * Mp + k], w, acc[n]);
      }
      const float b = bo[cl];
#pragma unroll
      for (int n = 0; n < NN; ++n)
        if (n < cnt) score_row(c, nodes[s_node[n]].out)[cl] = acc[n] + b;
    }
  }
}

}  // namespace n2nmn
