// K3: evaluation of one expression-tree node by one CTA, and the two executors built on it:
//   * tree_kernel : one CTA per question walks that question's nodes in Reverse-Polish order
//                   (every operand precedes its consumer), default;
//   * wave_kernel : one CTA per node of one tree depth across the whole batch — the
//                   depth-bucketed waves that replace TF Fold's dynamic batching
//                   (models_clevr/nmn3_model.py:49-159).
// Both call eval_node, so results are identical by construction.
//
// Find / Filter reach this kernel with their "find" map already in the arena (fused epilogue of
// the projection kernel), FindSameProperty with its projected feature map in mbuf.
#pragma once
#include "common.cuh"

namespace n2nmn {

struct NodeCtx {
  DevModel md;
  TextBufs tb;
  float* arena;        // [slots][HW]
  float* scores;       // [N][C]
  const float* mbuf;   // [mslots][HW][Mp]
};

// Shared-memory carve-up (floats); the host computes the same layout to size the launch.
struct NodeSmem {
  int HWp, pad, f, scratch, v, z, k, total;
};
__host__ __device__ inline NodeSmem node_smem_layout(int H, int W, int Dk, int Mp, int ksize,
                                                     int M) {
  NodeSmem s;
  const int HW = H * W;
  s.HWp = (HW + 3) & ~3;
  s.pad = ((H + ksize - 1) * (W + ksize - 1) + 3) & ~3;
  s.f = 2 * ((Dk + 3) & ~3);
  s.scratch = 1024;
  s.v = 3 * Mp;
  s.z = (2 * (HW + 2) + 3) & ~3;
  s.k = ksize * ksize * ((M + 127) / 128 * 128);
  s.total = 2 * s.HWp + s.pad + s.f + s.scratch + s.v + 64 + s.z + s.k;
  return s;
}

struct SmemPtrs {
  float *a0, *a1, *pad, *f, *scratch, *v0, *v1, *v2, *red, *z, *k;
};
__device__ __forceinline__ SmemPtrs carve(float* base, const DevModel& md) {
  const NodeSmem L = node_smem_layout(md.H, md.W, md.Dk, md.Mp, md.ksize, md.M);
  SmemPtrs s;
  s.a0 = base;
  s.a1 = s.a0 + L.HWp;
  s.pad = s.a1 + L.HWp;
  s.f = s.pad + L.pad;
  s.scratch = s.f + L.f;
  s.v0 = s.scratch + L.scratch;
  s.v1 = s.v0 + md.Mp;
  s.v2 = s.v1 + md.Mp;
  s.red = s.v2 + md.Mp;
  s.z = s.red + 64;
  s.k = s.z + L.z;
  return s;
}

// ---- building blocks ---------------------------------------------------------------------------
__device__ __forceinline__ void load_att(float* dst, const float* arena, int slot, int HW) {
  const float* src = arena + (size_t)slot * HW;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) dst[p] = src[p];
}

// tf.nn.softmax over the flattened map (models_clevr/nmn3_modules.py:170-172), in place.
__device__ __forceinline__ void softmax_inplace(float* a, int HW, float* red) {
  float mx = -INFINITY;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) mx = fmaxf(mx, a[p]);
  mx = block_reduce<1>(mx, red);
  float sum = 0.f;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    const float e = expf(a[p] - mx);
    a[p] = e;
    sum += e;
  }
  sum = block_reduce<0>(sum, red);
  for (int p = threadIdx.x; p < HW; p += blockDim.x) a[p] = a[p] / sum;
  __syncthreads();
}

// att_feat = reduce_sum(image_feat_grid * att_softmax, [1,2]) (nmn3_modules.py:174): f[d] for the
// node's image, read in place from the bound feature grid (the tf.gather copy of :49-51 is never
// materialised). float4 over channels; pixel range split across thread slices when Dk is small.
__device__ __forceinline__ void attention_pool(const DevModel& md, int b, const float* soft,
                                               float* f, float* scratch) {
  const int HW = md.HW, pitch = md.feat_pitch;
  const int ng = (md.Dk + 3) >> 2;
  const float4* X = reinterpret_cast<const float4*>(md.feat + (size_t)b * HW * pitch);
  const int pitch4 = pitch >> 2;
  const int nthreads = blockDim.x;
  if (ng >= nthreads) {
    for (int g = threadIdx.x; g < ng; g += nthreads) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int p = 0; p < HW; ++p) {
        const float4 x = __ldg(X + (size_t)p * pitch4 + g);
        const float s = soft[p];
        acc.x = fmaf(s, x.x, acc.x); acc.y = fmaf(s, x.y, acc.y);
        acc.z = fmaf(s, x.z, acc.z); acc.w = fmaf(s, x.w, acc.w);
      }
      reinterpret_cast<float4*>(f)[g] = acc;
    }
  } else {
    int slices = nthreads / ng;
    if (slices * ng * 4 > 1024) slices = 1024 / (ng * 4);
    const int g = threadIdx.x % ng, sl = threadIdx.x / ng;
    if (sl < slices) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int p = sl; p < HW; p += slices) {
        const float4 x = __ldg(X + (size_t)p * pitch4 + g);
        const float s = soft[p];
        acc.x = fmaf(s, x.x, acc.x); acc.y = fmaf(s, x.y, acc.y);
        acc.z = fmaf(s, x.z, acc.z); acc.w = fmaf(s, x.w, acc.w);
      }
      reinterpret_cast<float4*>(scratch)[sl * ng + g] = acc;
    }
    __syncthreads();
    for (int d = threadIdx.x; d < ng * 4; d += nthreads) {
      float s = 0.f;
      for (int k = 0; k < slices; ++k) s += scratch[k * ng * 4 + d];
      f[d] = s;
    }
  }
  __syncthreads();
}

// out[c] = bias[c] + Σ_k in[k]·W[k*M + c]  for c < M (zero for M <= c < Mp). fc_att etc.
__device__ __forceinline__ void gemv_cols(const float* in, int L, const float* __restrict__ W,
                                          const float* __restrict__ bias, float* out, int M,
                                          int Mp) {
  for (int c = threadIdx.x; c < Mp; c += blockDim.x) {
    float acc = 0.f;
    if (c < M) {
      acc = bias[c];
      const float* w = W + c;
      int k = 0;
      for (; k + 8 <= L; k += 8) {
        float wv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) wv[u] = __ldg(w + (size_t)(k + u) * M);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc = fmaf(in[k + u], wv[u], acc);
      }
      for (; k < L; ++k) acc = fmaf(in[k], __ldg(w + (size_t)k * M), acc);
    }
    out[c] = acc;
  }
  __syncthreads();
}

// scores[c] = b[c] + Σ_k z[k]·W[k*C + c]: the fc('fc_scores') / fc('fc_eltwise') heads
// (nmn3_modules.py:278,302,334,450,493). 8 thread groups split k, 32 lanes span c.
__device__ __forceinline__ void small_fc(const float* z, int L, const float* __restrict__ W,
                                         const float* __restrict__ bias, int C, float* out,
                                         float* scratch) {
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5, G = blockDim.x >> 5;
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int c = c0 + lane;
    float acc = 0.f;
    if (c < C)
      for (int k = g; k < L; k += G) acc = fmaf(z[k], __ldg(W + (size_t)k * C + c), acc);
    scratch[g * 32 + lane] = acc;
    __syncthreads();
    if (g == 0 && c < C) {
      float s = bias[c];
      for (int j = 0; j < G; ++j) s += scratch[j * 32 + lane];
      out[c] = s;
    }
    __syncthreads();
  }
}

// [a.flat, min, max] (Count / EqualNum...) or [min, mean, max] (Exist) into z.
__device__ __forceinline__ void minmax(const float* a, int HW, float* red, float& mn, float& mx,
                                       float& sum) {
  float lmn = INFINITY, lmx = -INFINITY, ls = 0.f;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    const float v = a[p];
    lmn = fminf(lmn, v); lmx = fmaxf(lmx, v); ls += v;
  }
  mn = block_reduce<2>(lmn, red);
  mx = block_reduce<1>(lmx, red);
  sum = block_reduce<0>(ls, red);
}

// ---- the modules -------------------------------------------------------------------------------
template <int KS>
__device__ __forceinline__ void eval_transform(const NodeCtx& c, const NodeRec& nd,
                                               const SmemPtrs& s) {
  // TransformModule, conv variant (models_clevr/nmn3_modules.py:185-216, SHAPES :71-101):
  // SAME cross-correlation of the 1-channel map with [KS,KS,1,M], ∘ text, l2norm over M, ·w2 + b2
  const DevModel& md = c.md;
  const int H = md.H, W = md.W, HW = md.HW, M = md.M;
  const int PW = W + KS - 1, PH = H + KS - 1, R = (KS - 1) / 2;
  const int Mq = (M + 127) / 128 * 128;
  for (int i = threadIdx.x; i < PH * PW; i += blockDim.x) s.pad[i] = 0.f;
  __syncthreads();
  const float* src = c.arena + (size_t)nd.in0 * HW;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    const int y = p / W, x = p - y * W;
    s.pad[(y + R) * PW + x + R] = src[p];
  }
  for (int i = threadIdx.x; i < KS * KS * Mq; i += blockDim.x) {
    const int tap = i / Mq, ch = i - tap * Mq;
    s.k[i] = (ch < M) ? md.conv_k[tap * M + ch] : 0.f;
  }
  const float* tau = c.tb.tau + (size_t)nd.text * md.Mp;
  for (int ch = threadIdx.x; ch < Mq; ch += blockDim.x) {
    const bool live = ch < M;
    s.v0[ch] = live ? tau[ch] : 0.f;
    s.v1[ch] = live ? md.elt_w[ES_TRANSFORM][ch] : 0.f;
    s.v2[ch] = live ? md.conv_b[ch] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float b2 = md.elt_b[ES_TRANSFORM][0];
  float* dst = c.arena + (size_t)nd.out * HW;
  for (int p = warp; p < HW; p += nwarps) {
    const int y = p / W, x = p - y * W;
    float win[KS * KS];
#pragma unroll
    for (int dy = 0; dy < KS; ++dy)
#pragma unroll
      for (int dx = 0; dx < KS; ++dx) win[dy * KS + dx] = s.pad[(y + dy) * PW + x + dx];
    float num = 0.f, den = 0.f;
    for (int c0 = lane * 4; c0 < Mq; c0 += 128) {
      float4 A = *reinterpret_cast<const float4*>(s.v2 + c0);
#pragma unroll
      for (int tap = 0; tap < KS * KS; ++tap) {
        const float4 k4 = *reinterpret_cast<const float4*>(s.k + tap * Mq + c0);
        A.x = fmaf(win[tap], k4.x, A.x); A.y = fmaf(win[tap], k4.y, A.y);
        A.z = fmaf(win[tap], k4.z, A.z); A.w = fmaf(win[tap], k4.w, A.w);
      }
      const float4 t4 = *reinterpret_cast<const float4*>(s.v0 + c0);
      const float4 w4 = *reinterpret_cast<const float4*>(s.v1 + c0);
      const float ex = A.x * t4.x, ey = A.y * t4.y, ez = A.z * t4.z, ew = A.w * t4.w;
      num = fmaf(ex, w4.x, num); num = fmaf(ey, w4.y, num);
      num = fmaf(ez, w4.z, num); num = fmaf(ew, w4.w, num);
      den = fmaf(ex, ex, den); den = fmaf(ey, ey, den);
      den = fmaf(ez, ez, den); den = fmaf(ew, ew, den);
    }
    num = warp_sum(num);
    den = warp_sum(den);
    if (lane == 0) dst[p] = num * rsqrtf(fmaxf(den, kEps)) + b2;
  }
}

__device__ __forceinline__ void eval_find_same_property(const NodeCtx& c, const NodeRec& nd,
                                                        const SmemPtrs& s) {
  // FindSamePropertyModule (models_clevr/nmn3_modules.py:134-183) and the VQA TransformModule
  // (models_vqa/nmn3_modules.py:123-171): l2norm_c(m ∘ τ ∘ φ)·w2 + b2 with φ = fc_att(pooled).
  const DevModel& md = c.md;
  const int HW = md.HW, Mp = md.Mp, M = md.M;
  load_att(s.a0, c.arena, nd.in0, HW);
  __syncthreads();
  softmax_inplace(s.a0, HW, s.red);
  attention_pool(md, nd.b, s.a0, s.f, s.scratch);
  gemv_cols(s.f, md.Dk, md.att_w[AS_FSP], md.att_b[AS_FSP], s.v0, M, Mp);
  const float* tauw = c.tb.tauw + (size_t)nd.text * Mp;   // τ∘w2
  const float* tau = c.tb.tau + (size_t)nd.text * Mp;
  for (int ch = threadIdx.x; ch < Mp; ch += blockDim.x) {
    const float phi = s.v0[ch];
    const float tp = tau[ch] * phi;
    s.v1[ch] = tauw[ch] * phi;   // coefficient of m in the numerator
    s.v2[ch] = tp * tp;          // coefficient of m² in the squared norm
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float b2 = md.elt_b[ES_FSP][0];
  const float* mimg = c.mbuf + (size_t)nd.aux * HW * Mp;
  float* dst = c.arena + (size_t)nd.out * HW;
  for (int p = warp; p < HW; p += nwarps) {
    const float4* mrow = reinterpret_cast<const float4*>(mimg + (size_t)p * Mp);
    float num = 0.f, den = 0.f;
    for (int q = lane; q < (Mp >> 2); q += 32) {
      const float4 m = __ldg(mrow + q);
      const float4 a = reinterpret_cast<const float4*>(s.v1)[q];
      const float4 d = reinterpret_cast<const float4*>(s.v2)[q];
      num = fmaf(m.x, a.x, num); num = fmaf(m.y, a.y, num);
      num = fmaf(m.z, a.z, num); num = fmaf(m.w, a.w, num);
      den = fmaf(m.x * m.x, d.x, den); den = fmaf(m.y * m.y, d.y, den);
      den = fmaf(m.z * m.z, d.z, den); den = fmaf(m.w * m.w, d.w, den);
    }
    num = warp_sum(num);
    den = warp_sum(den);
    if (lane == 0) dst[p] = num * rsqrtf(fmaxf(den, kEps)) + b2;
  }
}

__device__ __forceinline__ void normalize_and_score(const NodeCtx& c, const NodeRec& nd,
                                                    const SmemPtrs& s, float* e, int out_set) {
  // tf.nn.l2_normalize(e, 1) then fc('fc_eltwise') (nmn3_modules.py:448-450, 491-493)
  const DevModel& md = c.md;
  float ss = 0.f;
  for (int ch = threadIdx.x; ch < md.M; ch += blockDim.x) ss = fmaf(e[ch], e[ch], ss);
  ss = block_reduce<0>(ss, s.red);
  const float inv = rsqrtf(fmaxf(ss, kEps));
  for (int ch = threadIdx.x; ch < md.M; ch += blockDim.x) e[ch] *= inv;
  __syncthreads();
  small_fc(e, md.M, md.out_w[out_set], md.out_b[out_set], md.C,
           c.scores + (size_t)nd.out * md.C, s.scratch);
}

__device__ __forceinline__ void eval_describe(const NodeCtx& c, const NodeRec& nd,
                                              const SmemPtrs& s) {
  // DescribeModule (models_clevr/nmn3_modules.py:454-495, VQA models_vqa/nmn3_modules.py:193-240)
  const DevModel& md = c.md;
  load_att(s.a0, c.arena, nd.in0, md.HW);
  __syncthreads();
  softmax_inplace(s.a0, md.HW, s.red);
  attention_pool(md, nd.b, s.a0, s.f, s.scratch);
  gemv_cols(s.f, md.Dk, md.att_w[AS_DESCRIBE], md.att_b[AS_DESCRIBE], s.v0, md.M, md.Mp);
  const float* tau = c.tb.tau + (size_t)nd.text * md.Mp;
  for (int ch = threadIdx.x; ch < md.M; ch += blockDim.x) s.v1[ch] = tau[ch] * s.v0[ch];
  __syncthreads();
  normalize_and_score(c, nd, s, s.v1, OS_DESCRIBE);
}

__device__ __forceinline__ void eval_same_property(const NodeCtx& c, const NodeRec& nd,
                                                   const SmemPtrs& s) {
  // SamePropertyModule (models_clevr/nmn3_modules.py:402-452)
  const DevModel& md = c.md;
  load_att(s.a0, c.arena, nd.in0, md.HW);
  load_att(s.a1, c.arena, nd.in1, md.HW);
  __syncthreads();
  softmax_inplace(s.a0, md.HW, s.red);
  softmax_inplace(s.a1, md.HW, s.red);
  const int Dkp = (md.Dk + 3) & ~3;
  attention_pool(md, nd.b, s.a0, s.f, s.scratch);
  attention_pool(md, nd.b, s.a1, s.f + Dkp, s.scratch);
  gemv_cols(s.f, md.Dk, md.att_w[AS_SAMEPROP0], md.att_b[AS_SAMEPROP0], s.v0, md.M, md.Mp);
  gemv_cols(s.f + Dkp, md.Dk, md.att_w[AS_SAMEPROP1], md.att_b[AS_SAMEPROP1], s.v1, md.M, md.Mp);
  const float* tau = c.tb.tau + (size_t)nd.text * md.Mp;
  for (int ch = threadIdx.x; ch < md.M; ch += blockDim.x)
    s.v2[ch] = s.v0[ch] * tau[ch] * s.v1[ch];
  __syncthreads();
  normalize_and_score(c, nd, s, s.v2, OS_SAMEPROP);
}

__device__ __forceinline__ void eval_small_answer(const NodeCtx& c, const NodeRec& nd,
                                                  const SmemPtrs& s) {
  // Exist / Count / EqualNum / MoreNum / LessNum (models_clevr/nmn3_modules.py:258-400)
  const DevModel& md = c.md;
  const int HW = md.HW;
  float mn, mx, sum;
  int L, set;
  if (nd.op == OP_EXIST) {
    load_att(s.a0, c.arena, nd.in0, HW);
    __syncthreads();
    minmax(s.a0, HW, s.red, mn, mx, sum);
    if (threadIdx.x == 0) { s.z[0] = mn; s.z[1] = sum / (float)HW; s.z[2] = mx; }
    L = 3; set = SS_EXIST;
  } else {
    const int two = (nd.op != OP_COUNT);
    load_att(s.z, c.arena, nd.in0, HW);
    if (two) load_att(s.z + HW + 2, c.arena, nd.in1, HW);
    __syncthreads();
    minmax(s.z, HW, s.red, mn, mx, sum);
    if (threadIdx.x == 0) { s.z[HW] = mn; s.z[HW + 1] = mx; }
    if (two) {
      minmax(s.z + HW + 2, HW, s.red, mn, mx, sum);
      if (threadIdx.x == 0) { s.z[2 * HW + 2] = mn; s.z[2 * HW + 3] = mx; }
    }
    L = two ? 2 * (HW + 2) : HW + 2;
    set = (nd.op == OP_COUNT) ? SS_COUNT : (nd.op == OP_EQUAL_NUM) ? SS_EQUAL
        : (nd.op == OP_MORE_NUM) ? SS_MORE : SS_LESS;
  }
  __syncthreads();
  small_fc(s.z, L, md.sc_w[set], md.sc_b[set], md.C, c.scores + (size_t)nd.out * md.C,
           s.scratch);
}

template <int KS>
__device__ __forceinline__ void eval_node(const NodeCtx& c, const NodeRec& nd, const SmemPtrs& s) {
  const int HW = c.md.HW;
  switch (nd.op) {
    case OP_SCENE: {   // models_clevr/nmn3_modules.py:60-72; aux carries pos_val's bits
      float* dst = c.arena + (size_t)nd.out * HW;
      const float v = __int_as_float(nd.aux);
      for (int p = threadIdx.x; p < HW; p += blockDim.x) dst[p] = v;
      break;
    }
    case OP_FIND:      // already written by the projection kernel's epilogue
      break;
    case OP_FILTER: {  // min(input_0, Find(t,b)) (nmn3_modules.py:129-130); find part is in `out`
      float* dst = c.arena + (size_t)nd.out * HW;
      const float* a = c.arena + (size_t)nd.in0 * HW;
      for (int p = threadIdx.x; p < HW; p += blockDim.x) dst[p] = fminf(a[p], dst[p]);
      break;
    }
    case OP_AND:       // tf.minimum / tf.maximum (nmn3_modules.py:233,253)
    case OP_OR: {
      float* dst = c.arena + (size_t)nd.out * HW;
      const float* a = c.arena + (size_t)nd.in0 * HW;
      const float* b = c.arena + (size_t)nd.in1 * HW;
      for (int p = threadIdx.x; p < HW; p += blockDim.x)
        dst[p] = (nd.op == OP_AND) ? fminf(a[p], b[p]) : fmaxf(a[p], b[p]);
      break;
    }
    case OP_TRANSFORM: eval_transform<KS>(c, nd, s); break;
    case OP_FIND_SAME_PROPERTY: eval_find_same_property(c, nd, s); break;
    case OP_DESCRIBE: eval_describe(c, nd, s); break;
    case OP_SAME_PROPERTY: eval_same_property(c, nd, s); break;
    default: eval_small_answer(c, nd, s); break;
  }
}

// One CTA per question; q_ptr delimits the question's nodes (Reverse-Polish order) in `nodes`.
template <int KS>
__global__ void __launch_bounds__(kNodeThreads)
tree_kernel(const NodeCtx c, const NodeRec* __restrict__ nodes, const int32_t* __restrict__ q_ptr) {
  extern __shared__ float node_smem[];
  const SmemPtrs s = carve(node_smem, c.md);
  const int q = blockIdx.x;
  const int beg = q_ptr[q], end = q_ptr[q + 1];
  if (beg == end) {   // invalid layout: zeros(num_choices) (models_clevr/nmn3_model.py:144-155)
    for (int i = threadIdx.x; i < c.md.C; i += blockDim.x) c.scores[(size_t)q * c.md.C + i] = 0.f;
    return;
  }
  for (int i = beg; i < end; ++i) {
    const NodeRec nd = nodes[i];
    eval_node<KS>(c, nd, s);
    __syncthreads();   // arena writes of this node are visible to the block's next node
  }
}

// One CTA per node of one wave.
template <int KS>
__global__ void __launch_bounds__(kNodeThreads)
wave_kernel(const NodeCtx c, const NodeRec* __restrict__ nodes,
            const int32_t* __restrict__ wave_nodes, int first) {
  extern __shared__ float node_smem[];
  const SmemPtrs s = carve(node_smem, c.md);
  const NodeRec nd = nodes[wave_nodes[first + blockIdx.x]];
  eval_node<KS>(c, nd, s);
}

}  // namespace n2nmn
