// K3 (default executor): one thread-block cluster per question walks the question's nodes in
// Reverse-Polish order, keeping every live attention map of the question in SHARED memory.
//
// Why it looks like this: at batch 64 a question is a chain of 2-6 tiny dependent steps; what
// bounds it is latency (L2 round trips, barriers), not bandwidth or arithmetic. So
//   * the attention "stack" lives in smem, replicated in every CTA of the cluster — the cheap
//     modules (Scene / Filter / And / Or) are recomputed redundantly by every CTA and need no
//     cluster barrier and no global memory at all;
//   * parameters (conv filter bank, the root's answer-head weights) and the Find maps produced
//     by the projection kernel are fetched with cp.async in the prologue, behind PDL;
//   * only the heavy modules are split over the cluster: Transform / FindSameProperty by pixel,
//     the attention-pooled fc_att (Describe / SameProperty / FindSameProperty) by rows of the
//     stored map; pieces are exchanged through double-buffered distributed shared memory with ONE
//     cluster barrier per exchange.
// Results are identical to wave_kernel / eval_node (node_eval.cuh) up to fp32 summation order.
#pragma once
#include "node_eval.cuh"

namespace n2nmn {

constexpr int kTreeFindSlots = 10;   // Find / Filter maps prefetched per question
constexpr int kTransformPB = 5;      // pixels register-blocked per warp in the stencil
constexpr int kTreeNodeCap = 48;     // node records of one question kept in smem
constexpr int kTreeTextCap = 6;      // text vectors (tau, tau∘w2) of one question prefetched

struct TreeSmem {
  int HWp, stack, ftmp, outbuf, pad, v, part, z, k, head, nodes, tvec, total;
};
__host__ __device__ inline TreeSmem tree_smem_layout(int H, int W, int Mp, int ksize, int M, int C,
                                                     int stack_slots) {
  TreeSmem s;
  const int HW = H * W;
  s.HWp = (HW + 3) & ~3;
  s.stack = stack_slots * s.HWp;
  s.ftmp = kTreeFindSlots * s.HWp;
  s.outbuf = 2 * s.HWp;
  s.pad = ((H + ksize - 1) * (W + ksize - 1) + 3) & ~3;
  s.v = 3 * Mp;
  s.part = 4 * Mp;
  s.z = (2 * (HW + 2) + 3) & ~3;
  s.k = ksize * ksize * Mp;
  const int rows = (2 * (HW + 2) > M) ? 2 * (HW + 2) : M;
  s.head = (rows * C <= kHeadCapFloats) ? ((rows * C + 3) & ~3) : 0;
  s.nodes = kTreeNodeCap * (int)(sizeof(NodeRec) / sizeof(float));
  s.tvec = (Mp <= 512) ? kTreeTextCap * 2 * Mp : 0;
  s.total = s.stack + s.ftmp + s.outbuf + 2 * s.HWp + s.pad + kNodeScratch + s.v + s.part + 64 +
            s.z + s.k + s.head + s.nodes + s.tvec + 2 * Mp;
  return s;
}

struct TreePtrs {
  float *stack, *ftmp, *outbuf, *a0, *a1, *pad, *scratch, *v0, *v1, *v2, *part, *red, *z, *k, *head;
  int HWp, Mp;
};

// softmax of one map by a single warp (HW is a few hundred at most), in place.
__device__ __forceinline__ void warp_softmax(float* a, int HW) {
  const int lane = threadIdx.x & 31;
  float mx = -INFINITY;
  for (int p = lane; p < HW; p += 32) mx = fmaxf(mx, a[p]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int p = lane; p < HW; p += 32) {
    const float e = expf(a[p] - mx);
    a[p] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  for (int p = lane; p < HW; p += 32) a[p] = a[p] / sum;
}

// min / max / sum of one map by a single warp.
__device__ __forceinline__ void warp_minmax(const float* a, int HW, float& mn, float& mx,
                                            float& sum) {
  const int lane = threadIdx.x & 31;
  float lmn = INFINITY, lmx = -INFINITY, ls = 0.f;
  for (int p = lane; p < HW; p += 32) {
    const float v = a[p];
    lmn = fminf(lmn, v); lmx = fmaxf(lmx, v); ls += v;
  }
  mn = warp_min(lmn); mx = warp_max(lmx); sum = warp_sum(ls);
}

// After a pixel-split phase and its cluster barrier: assemble the full map in the local stack.
// Work unit u (a pixel, or a horizontal run of `run` pixels) was done by global warp u % gwarps.
__device__ __forceinline__ void gather_pixels(const Coop& co, const float* outbuf, float* dst,
                                              int HW, int W, int run) {
  const int nwarps = blockDim.x >> 5, gwarps = co.size * nwarps;
  const int runs_x = (W + run - 1) / run;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    const int y = p / W, x = p - y * W;
    const int unit = (run == 1) ? p : y * runs_x + x / run;
    dst[p] = co.peer(outbuf, (unit % gwarps) / nwarps)[p];
  }
}

template <int KS>
__global__ void __launch_bounds__(kNodeThreads, 2)
tree_kernel(const NodeCtx c, const NodeRec* __restrict__ nodes, const int32_t* __restrict__ q_ptr,
            int csize, int stack_slots, int write_arena) {
  extern __shared__ __align__(16) float tree_smem[];
  const DevModel& md = c.md;
  const int HW = md.HW, Mp = md.Mp, M = md.M;
  const TreeSmem L = tree_smem_layout(md.H, md.W, Mp, md.ksize, M, md.C, stack_slots);
  TreePtrs s;
  s.HWp = L.HWp; s.Mp = Mp;
  s.stack = tree_smem;
  s.ftmp = s.stack + L.stack;
  s.outbuf = s.ftmp + L.ftmp;
  s.a0 = s.outbuf + L.outbuf;
  s.a1 = s.a0 + L.HWp;
  s.pad = s.a1 + L.HWp;
  s.scratch = s.pad + L.pad;
  s.v0 = s.scratch + kNodeScratch;
  s.v1 = s.v0 + Mp;
  s.v2 = s.v1 + Mp;
  s.part = s.v2 + Mp;
  s.red = s.part + L.part;
  s.z = s.red + 64;
  s.k = s.z + L.z;
  s.head = L.head ? s.k + L.k : nullptr;
  NodeRec* s_nodes = reinterpret_cast<NodeRec*>(s.k + L.k + L.head);
  float* s_tvec = reinterpret_cast<float*>(s_nodes) + L.nodes;    // [kTreeTextCap][2][Mp]
  float* s_tw2 = s_tvec + L.tvec;                                   // Transform conv_eltwise w
  float* s_tcb = s_tw2 + Mp;                                        // Transform conv bias

  Coop co;
  co.size = csize;
  co.rank = (csize > 1) ? (int)cg::this_cluster().block_rank() : 0;
  const int q = blockIdx.x / csize;
  const int beg = q_ptr[q], end = q_ptr[q + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int gwarp = co.rank * nwarps + warp, gwarps = co.size * nwarps;

  if (threadIdx.x == 0) N2NMN_STAMP(2, 0);
  // ---- prologue part 1: parameters (weights + launch tables only; overlaps the predecessors)
  const float* head_w = nullptr;
  const bool nodes_in_smem = (end - beg) <= kTreeNodeCap;
  if (beg < end) {
    if (nodes_in_smem) {   // node records: the loop below must not wait on L2 after pdl_wait
      const int nwords = (end - beg) * (int)(sizeof(NodeRec) / 4);
      const int32_t* src = reinterpret_cast<const int32_t*>(nodes + beg);
      int32_t* dst = reinterpret_cast<int32_t*>(s_nodes);
      for (int j = threadIdx.x; j < nwords; j += blockDim.x) dst[j] = src[j];
    }
    bool has_transform = false;
    for (int i = beg; i < end; ++i) has_transform |= (nodes[i].op == OP_TRANSFORM);
    if (has_transform) {
      stage_async(s.k, md.conv_k, KS * KS * Mp);
      for (int ch = threadIdx.x; ch < Mp; ch += blockDim.x) {
        s_tw2[ch] = (ch < M) ? md.elt_w[ES_TRANSFORM][ch] : 0.f;
        s_tcb[ch] = (ch < M) ? md.conv_b[ch] : 0.f;
      }
    }
    if (co.rank == 0 && s.head != nullptr) {
      const int rop = nodes[end - 1].op;
      const float* w = nullptr;
      int rows = 0;
      switch (rop) {
        case OP_EXIST: w = md.sc_w[SS_EXIST]; rows = 3; break;
        case OP_COUNT: w = md.sc_w[SS_COUNT]; rows = HW + 2; break;
        case OP_EQUAL_NUM: w = md.sc_w[SS_EQUAL]; rows = 2 * (HW + 2); break;
        case OP_MORE_NUM: w = md.sc_w[SS_MORE]; rows = 2 * (HW + 2); break;
        case OP_LESS_NUM: w = md.sc_w[SS_LESS]; rows = 2 * (HW + 2); break;
        case OP_SAME_PROPERTY: w = md.out_w[OS_SAMEPROP]; rows = M; break;
        case OP_DESCRIBE: w = md.out_w[OS_DESCRIBE]; rows = M; break;
        default: break;
      }
      if (w) { stage_async(s.head, w, rows * md.C); head_w = s.head; }
    }
  }
  // ---- the attention arena, stored maps and text projections come from the preceding kernels
  if (threadIdx.x == 0) N2NMN_STAMP(2, 1);
  pdl_wait();
  if (threadIdx.x == 0) N2NMN_STAMP(2, 2);
  if (beg == end) {   // invalid layout: zeros(num_choices) (models_clevr/nmn3_model.py:144-155)
    if (co.rank == 0)
      for (int i = threadIdx.x; i < md.C; i += blockDim.x) c.scores[(size_t)q * md.C + i] = 0.f;
    return;
  }
  // ---- prologue part 2: the question's Find maps (written by the projection kernel's epilogue)
  //      and the text vectors of its Transform / pooled nodes (written by the text kernel)
  const NodeRec* qnodes = nodes_in_smem ? s_nodes : nodes + beg;
  {
    int nf = 0, nt = 0;
    for (int i = beg; i < end; ++i) {
      const int op = qnodes[i - beg].op;
      if (L.tvec && nt < kTreeTextCap &&
          (op == OP_TRANSFORM || op == OP_FIND_SAME_PROPERTY || op == OP_DESCRIBE ||
           op == OP_SAME_PROPERTY)) {
        const size_t row = (size_t)qnodes[i - beg].text * Mp;
        stage_async(s_tvec + (nt * 2) * Mp, c.tb.tau + row, Mp);
        stage_async(s_tvec + (nt * 2 + 1) * Mp, c.tb.tauw + row, Mp);
        ++nt;
      }
      if (op == OP_FIND || op == OP_FILTER) {   // staged apart: stack slots are reused over time
        if (nf < kTreeFindSlots)
          stage_async(s.ftmp + nf * L.HWp, c.arena + (size_t)nodes[i].out * HW, HW);
        ++nf;
      }
    }
    cp_async_commit_wait_all();
    __syncthreads();
  }

  if (threadIdx.x == 0) N2NMN_STAMP(2, 3);
  int exch = 0;        // cluster exchanges so far: selects the double-buffered outbuf / part half
  int nfilter = 0, ntext = 0;
  for (int i = beg; i < end; ++i) {
    const NodeRec nd = qnodes[i - beg];
    // text vectors of this node: prefetched copy if it got a slot, else straight from L2
    const float* tau = nullptr;
    const float* tauw = nullptr;
    if (nd.op == OP_TRANSFORM || nd.op == OP_FIND_SAME_PROPERTY || nd.op == OP_DESCRIBE ||
        nd.op == OP_SAME_PROPERTY) {
      if (L.tvec && ntext < kTreeTextCap) {
        tau = s_tvec + (ntext * 2) * Mp;
        tauw = tau + Mp;
      } else {
        tau = c.tb.tau + (size_t)nd.text * Mp;
        tauw = c.tb.tauw + (size_t)nd.text * Mp;
      }
      ++ntext;
    }
    float* out = (nd.so >= 0) ? s.stack + nd.so * L.HWp : nullptr;
    const float* in0 = (nd.s0 >= 0) ? s.stack + nd.s0 * L.HWp : nullptr;
    const float* in1 = (nd.s1 >= 0) ? s.stack + nd.s1 * L.HWp : nullptr;
    switch (nd.op) {
      case OP_FIND: {        // computed by the projection kernel's epilogue; prefetched above
        const float* f = (nfilter < kTreeFindSlots) ? s.ftmp + nfilter * L.HWp
                                                    : c.arena + (size_t)nd.out * HW;
        ++nfilter;
        for (int p = threadIdx.x; p < HW; p += blockDim.x) out[p] = f[p];
        break;
      }
      case OP_SCENE: {       // models_clevr/nmn3_modules.py:60-72
        const float v = __int_as_float(nd.aux);
        for (int p = threadIdx.x; p < HW; p += blockDim.x) out[p] = v;
        break;
      }
      case OP_FILTER: {      // min(input_0, Find(t,b)) (nmn3_modules.py:129-130)
        const float* f = (nfilter < kTreeFindSlots) ? s.ftmp + nfilter * L.HWp
                                                    : c.arena + (size_t)nd.out * HW;
        ++nfilter;
        for (int p = threadIdx.x; p < HW; p += blockDim.x) out[p] = fminf(in0[p], f[p]);
        break;
      }
      case OP_AND:           // tf.minimum / tf.maximum (nmn3_modules.py:233,253)
      case OP_OR:
        for (int p = threadIdx.x; p < HW; p += blockDim.x)
          out[p] = (nd.op == OP_AND) ? fminf(in0[p], in1[p]) : fmaxf(in0[p], in1[p]);
        break;
      case OP_TRANSFORM: {
        if (threadIdx.x == 0) N2NMN_STAMP(2, 20);
        // TransformModule, conv variant (models_clevr/nmn3_modules.py:185-216, SHAPES :71-101)
        const int Hh = md.H, Ww = md.W;
        const int PW = Ww + KS - 1, PH = Hh + KS - 1, R = (KS - 1) / 2;
        for (int j = threadIdx.x; j < PH * PW; j += blockDim.x) s.pad[j] = 0.f;
        for (int ch = threadIdx.x; ch < Mp; ch += blockDim.x) {
          s.v0[ch] = (ch < M) ? tau[ch] : 0.f;
          s.v1[ch] = s_tw2[ch];
          s.v2[ch] = s_tcb[ch];
        }
        __syncthreads();
        for (int p = threadIdx.x; p < HW; p += blockDim.x) {
          const int y = p / Ww, x = p - y * Ww;
          s.pad[(y + R) * PW + x + R] = in0[p];
        }
        cp_async_commit_wait_all();   // the filter bank (prologue)
        __syncthreads();
        if (threadIdx.x == 0) N2NMN_STAMP(2, 21);
        const float b2 = md.elt_b[ES_TRANSFORM][0];
        float* ob = s.outbuf + (exch & 1) * L.HWp;
        // Each warp owns horizontal runs of kTransformPB pixels: their stencil windows overlap, so
        // one (KS x (PB+KS-1)) window is read into registers once and every filter-bank read is
        // shared by the PB pixels (the stencil is shared-memory-bandwidth bound otherwise).
        constexpr int WW = kTransformPB + KS - 1;
        const int runs_x = (Ww + kTransformPB - 1) / kTransformPB;
        for (int run = gwarp; run < Hh * runs_x; run += gwarps) {
          const int y = run / runs_x, x0 = (run - y * runs_x) * kTransformPB;
          float win[KS][WW];
#pragma unroll
          for (int dy = 0; dy < KS; ++dy)
#pragma unroll
            for (int i = 0; i < WW; ++i)
              win[dy][i] = s.pad[(y + dy) * PW + min(x0 + i, PW - 1)];
          float num[kTransformPB], den[kTransformPB];
#pragma unroll
          for (int j = 0; j < kTransformPB; ++j) { num[j] = 0.f; den[j] = 0.f; }
          for (int c0 = lane * 4; c0 < Mp; c0 += 128) {
            // 4 channels x PB pixels of accumulators, updated with two-wide fp32 FMAs (FFMA2):
            // the stencil is FMA-issue bound (0.94 MFMA per node)
            float2 Alo[kTransformPB], Ahi[kTransformPB];
            const float4 bias4 = *reinterpret_cast<const float4*>(s.v2 + c0);
#pragma unroll
            for (int j = 0; j < kTransformPB; ++j) {
              Alo[j] = make_float2(bias4.x, bias4.y);
              Ahi[j] = make_float2(bias4.z, bias4.w);
            }
#pragma unroll
            for (int dy = 0; dy < KS; ++dy) {
#pragma unroll
              for (int dx = 0; dx < KS; ++dx) {
                const float4 k4 =
                    *reinterpret_cast<const float4*>(s.k + (dy * KS + dx) * Mp + c0);
                const float2 klo = make_float2(k4.x, k4.y), khi = make_float2(k4.z, k4.w);
#pragma unroll
                for (int j = 0; j < kTransformPB; ++j) {
                  const float wv = win[dy][dx + j];
                  const float2 w2v = make_float2(wv, wv);
                  Alo[j] = __ffma2_rn(w2v, klo, Alo[j]);
                  Ahi[j] = __ffma2_rn(w2v, khi, Ahi[j]);
                }
              }
            }
            const float4 t4 = *reinterpret_cast<const float4*>(s.v0 + c0);
            const float4 w4 = *reinterpret_cast<const float4*>(s.v1 + c0);
#pragma unroll
            for (int j = 0; j < kTransformPB; ++j) {
              const float ex = Alo[j].x * t4.x, ey = Alo[j].y * t4.y, ez = Ahi[j].x * t4.z,
                          ew = Ahi[j].y * t4.w;
              num[j] = fmaf(ex, w4.x, num[j]); num[j] = fmaf(ey, w4.y, num[j]);
              num[j] = fmaf(ez, w4.z, num[j]); num[j] = fmaf(ew, w4.w, num[j]);
              den[j] = fmaf(ex, ex, den[j]); den[j] = fmaf(ey, ey, den[j]);
              den[j] = fmaf(ez, ez, den[j]); den[j] = fmaf(ew, ew, den[j]);
            }
          }
#pragma unroll
          for (int j = 0; j < kTransformPB; ++j) {
            const float n = warp_sum(num[j]), d = warp_sum(den[j]);
            if (lane == 0 && x0 + j < Ww) ob[y * Ww + x0 + j] = n * rsqrtf(fmaxf(d, kEps)) + b2;
          }
        }
        if (threadIdx.x == 0) N2NMN_STAMP(2, 22);
        co.sync();
        if (threadIdx.x == 0) N2NMN_STAMP(2, 23);
        gather_pixels(co, ob, out, HW, Ww, kTransformPB);
        if (threadIdx.x == 0) N2NMN_STAMP(2, 24);
        ++exch;
        break;
      }
      case OP_FIND_SAME_PROPERTY: {
        // FindSamePropertyModule (models_clevr/nmn3_modules.py:134-183) / VQA TransformModule
        // (models_vqa/nmn3_modules.py:123-171): l2norm_c(m ∘ τ ∘ φ)·w2 + b2, φ = Σ_p s_p·G[p,:]
        for (int p = threadIdx.x; p < HW; p += blockDim.x) s.a0[p] = in0[p];
        __syncthreads();
        if (warp == 0) warp_softmax(s.a0, HW);
        __syncthreads();
        float* part = s.part + (exch & 1) * 2 * Mp;
        int p0, p1;
        coop_range(co, HW, p0, p1);
        gemv_partial(s.a0 + p0, p0, p1, c.mbuf + (size_t)nd.aux2 * HW * Mp, Mp, part, s.scratch);
        co.sync();
        ++exch;
        sum_partials(co, part, nullptr, s.v0, M, Mp);
        for (int ch = threadIdx.x; ch < Mp; ch += blockDim.x) {
          const float phi = s.v0[ch];
          const float tp = tau[ch] * phi;
          s.v1[ch] = tauw[ch] * phi;
          s.v2[ch] = tp * tp;
        }
        __syncthreads();
        const float b2 = md.elt_b[ES_FSP][0];
        const float* mimg = c.mbuf + (size_t)nd.aux * HW * Mp;
        float* ob = s.outbuf + (exch & 1) * L.HWp;
        for (int p = gwarp; p < HW; p += gwarps) {
          const float4* mrow = reinterpret_cast<const float4*>(mimg + (size_t)p * Mp);
          float num = 0.f, den = 0.f;
          for (int qd = lane; qd < (Mp >> 2); qd += 32) {
            const float4 m = __ldg(mrow + qd);
            const float4 a = reinterpret_cast<const float4*>(s.v1)[qd];
            const float4 d = reinterpret_cast<const float4*>(s.v2)[qd];
            num = fmaf(m.x, a.x, num); num = fmaf(m.y, a.y, num);
            num = fmaf(m.z, a.z, num); num = fmaf(m.w, a.w, num);
            den = fmaf(m.x * m.x, d.x, den); den = fmaf(m.y * m.y, d.y, den);
            den = fmaf(m.z * m.z, d.z, den); den = fmaf(m.w * m.w, d.w, den);
          }
          num = warp_sum(num);
          den = warp_sum(den);
          if (lane == 0) ob[p] = num * rsqrtf(fmaxf(den, kEps)) + b2;
        }
        co.sync();
        gather_pixels(co, ob, out, HW, md.W, 1);
        ++exch;
        break;
      }
      case OP_DESCRIBE:
      case OP_SAME_PROPERTY: {
        // DescribeModule (nmn3_modules.py:454-495) / SamePropertyModule (:402-452)
        const bool two = (nd.op == OP_SAME_PROPERTY);
        if (threadIdx.x == 0) N2NMN_STAMP(2, 26);
        for (int p = threadIdx.x; p < HW; p += blockDim.x) {
          s.a0[p] = in0[p];
          if (two) s.a1[p] = in1[p];
        }
        __syncthreads();
        if (warp == 0) warp_softmax(s.a0, HW);
        if (two && warp == 1) warp_softmax(s.a1, HW);
        __syncthreads();
        if (threadIdx.x == 0) N2NMN_STAMP(2, 27);
        float* part = s.part + (exch & 1) * 2 * Mp;
        int p0, p1;
        coop_range(co, HW, p0, p1);
        const size_t map_floats = (size_t)HW * Mp;
        gemv_partial(s.a0 + p0, p0, p1, c.mbuf + nd.aux * map_floats, Mp, part, s.scratch);
        if (two)
          gemv_partial(s.a1 + p0, p0, p1, c.mbuf + nd.aux2 * map_floats, Mp, part + Mp, s.scratch);
        if (threadIdx.x == 0) N2NMN_STAMP(2, 28);
        co.sync();
        if (threadIdx.x == 0) N2NMN_STAMP(2, 29);
        ++exch;
        if (co.rank == 0) {   // the tail is tiny: one CTA finishes it
          sum_partials(co, part, nullptr, s.v0, M, Mp);
          if (two) sum_partials(co, part + Mp, nullptr, s.v1, M, Mp);
          float ss = 0.f;
          for (int ch = threadIdx.x; ch < M; ch += blockDim.x) {
            const float e = two ? s.v0[ch] * tau[ch] * s.v1[ch] : tau[ch] * s.v0[ch];
            s.v2[ch] = e;
            ss = fmaf(e, e, ss);
          }
          ss = block_reduce<0>(ss, s.red);
          const float inv = rsqrtf(fmaxf(ss, kEps));   // tf.nn.l2_normalize(e, 1)
          for (int ch = threadIdx.x; ch < M; ch += blockDim.x) s.v2[ch] *= inv;
          cp_async_commit_wait_all();
          __syncthreads();
          if (threadIdx.x == 0) N2NMN_STAMP(2, 30);
          const int os = two ? OS_SAMEPROP : OS_DESCRIBE;
          small_fc(s.v2, M, head_w ? head_w : md.out_w[os], md.out_b[os], md.C,
                   c.scores + (size_t)nd.out * md.C, s.scratch);
        }
        break;
      }
      default: {
        // Exist / Count / EqualNum / MoreNum / LessNum (nmn3_modules.py:258-400): rank 0 only,
        // everything it needs is in its local stack
        if (co.rank != 0) break;
        int Lz, set;
        if (nd.op == OP_EXIST) {
          if (warp == 0) {
            float mn, mx, sum;
            warp_minmax(in0, HW, mn, mx, sum);
            if (lane == 0) { s.z[0] = mn; s.z[1] = sum / (float)HW; s.z[2] = mx; }
          }
          Lz = 3; set = SS_EXIST;
        } else {
          const bool two = (nd.op != OP_COUNT);
          for (int p = threadIdx.x; p < HW; p += blockDim.x) {
            s.z[p] = in0[p];
            if (two) s.z[HW + 2 + p] = in1[p];
          }
          if (warp == 0) {
            float mn, mx, sum;
            warp_minmax(in0, HW, mn, mx, sum);
            if (lane == 0) { s.z[HW] = mn; s.z[HW + 1] = mx; }
          } else if (two && warp == 1) {
            float mn, mx, sum;
            warp_minmax(in1, HW, mn, mx, sum);
            if (lane == 0) { s.z[2 * HW + 2] = mn; s.z[2 * HW + 3] = mx; }
          }
          Lz = two ? 2 * (HW + 2) : HW + 2;
          set = (nd.op == OP_COUNT) ? SS_COUNT : (nd.op == OP_EQUAL_NUM) ? SS_EQUAL
              : (nd.op == OP_MORE_NUM) ? SS_MORE : SS_LESS;
        }
        cp_async_commit_wait_all();
        __syncthreads();
        small_fc(s.z, Lz, head_w ? head_w : md.sc_w[set], md.sc_b[set], md.C,
                 c.scores + (size_t)nd.out * md.C, s.scratch);
        break;
      }
    }
    __syncthreads();   // this node's stack writes are visible to the CTA's next node
    if (threadIdx.x == 0) N2NMN_STAMP(2, 4 + (i - beg));
    if (write_arena && co.rank == 0 && out != nullptr && nd.op != OP_FIND) {
      float* g = c.arena + (size_t)nd.out * HW;
      for (int p = threadIdx.x; p < HW; p += blockDim.x) g[p] = out[p];
    }
  }
  // peers may still be reading this CTA's outbuf / part through distributed shared memory
  if (csize > 1) cg::this_cluster().sync();
}

}  // namespace n2nmn
