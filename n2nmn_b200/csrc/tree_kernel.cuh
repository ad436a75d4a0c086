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
//     cluster barrier per exchange;
//   * the Transform stencil (the one arithmetic-heavy step) runs on TF32 mma.sync fragments; the
//     exact-fp32 CUDA-core stencil is kept for the verification mode (kTreeFp32Stencil).
// With many batches in flight the pool launches ONE CTA per question instead (no cluster
// barriers at all): less latency hiding per question, more questions per second.
// Results agree with wave_kernel / eval_node (node_eval.cuh) up to fp32 summation order and the
// TF32 rounding of the stencil operands.
#pragma once
#include "node_eval.cuh"

namespace n2nmn {

constexpr int kTreeWriteArena = 1;     // launch flags
constexpr int kTreeFp32Stencil = 2;    // exact-fp32 Transform stencil on the CUDA cores (verification)
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
  s.pad = (2 * (H + ksize - 1) * (W + ksize - 1) + 3) & ~3;   // padded map + zero guard copy
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

// ---- Transform stencil on the tensor cores ---------------------------------------------------------
// conv_maps of TransformModule (models_clevr/nmn3_modules.py:197-201, SHAPES :83-87) is the GEMM
//   maps[p, c] = Σ_tap window(p)[tap] · K[tap, c]        (H·W x KS² x M, 0.94 MFMA per node)
// and its consumer (∘τ, l2-normalise over c, ·w2) is a per-row reduction over c. The CUDA cores
// need ~5.6-11 K cycles per node for it; as m16n8k8 TF32 MMAs with the window gathered straight out
// of the zero-padded attention map it is ~1.5 K. Register-fragment `mma.sync` rather than tcgen05:
// the product is 160 x 256 x 32 per node, issued from inside a latency-bound walk, and its result
// is consumed by the issuing lanes' own row reductions — no TMEM round trip, no descriptors.
// Operands are rounded to TF32 once, where they are written to shared memory (the padded map when
// it is filled, the filter bank after it has been staged): cvt.rna.tf32 is a three-instruction
// sequence on sm_100 and the fragment loads are the inner loop. Accumulation is fp32.
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void mma_tf32_m16n8k8(float (&d)[4], const uint32_t (&a)[4],
                                                 const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
constexpr int kStencilRowsPerPass = 128;   // rows reduced through `scratch` at a time (8 m-tiles)

// Owner rank of pixel p when the 16-row tiles of the map are dealt to the cluster in contiguous
// blocks (the tensor-core stencil's split).
__device__ __forceinline__ int stencil_owner(int p, int HW, int csize) {
  const int n_mt = (HW + 15) >> 4, per = (n_mt + csize - 1) / csize;
  return (p >> 4) / per;
}

// One Transform node on this CTA's share of the pixels. pad: zero-padded input map
// [(H+KS-1)][(W+KS-1)], TF32-rounded, followed by an all-zero copy of the same size; kbank: [KS*KS][Mp], TF32-rounded;
// tau / w2 / cbias: [Mp] (zero beyond M);
// scratch: >= 8 warps x 128 rows x 2 floats; ob: this CTA's output pixels (indexed by p).
template <int KS>
__device__ __forceinline__ void transform_stencil_mma(const Coop& co, const float* pad,
                                                      const float* kbank, const float* tau,
                                                      const float* w2, const float* cbias,
                                                      float b2, float* scratch, float* ob,
                                                      int Hh, int Ww, int Mp) {
  constexpr int KK = KS * KS, KSTEPS = (KK + 7) / 8;
  const int HW = Hh * Ww, PW = Ww + KS - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int n_mt = (HW + 15) >> 4, per = (n_mt + co.size - 1) / co.size;
  const int mt_beg = co.rank * per, mt_end = min(n_mt, mt_beg + per);
  // taps beyond KS² read the all-zero guard copy that follows the padded map (any row base lands in
  // it); rows beyond the map use base 0 and are never written out. No clamps in the inner loop.
  const int zero_off = (Hh + KS - 1) * PW;
  const uint32_t inv_w = (1u << 20) / (uint32_t)Ww + 1u;   // p / Ww for p < 4096
  // window offsets of the taps this lane feeds: 8*ks + t and 8*ks + t + 4
  int toff[KSTEPS][2];
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int tap = 8 * ks + t + 4 * h;
      toff[ks][h] = (tap < KK) ? (tap / KS) * PW + (tap % KS) : zero_off;   // -> the zero guard
    }
  for (int pass = mt_beg; pass < mt_end; pass += kStencilRowsPerPass / 16) {
    const int pass_end = min(mt_end, pass + kStencilRowsPerPass / 16);
    bool first_cb = true;
    for (int cb = warp * 32; cb < Mp; cb += nwarps * 32) {   // this warp's 32-channel blocks
      uint32_t bf[KSTEPS][4][2];
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int tap = 8 * ks + t + 4 * h;
            const uint32_t v = __float_as_uint(kbank[min(tap, KK - 1) * Mp + cb + 8 * j + g]);
            bf[ks][j][h] = (tap < KK) ? v : 0u;
          }
      // per-channel constants of this lane's 8 output columns (c = cb + 8j + 2t, +1)
      float2 cb2[4], tw2[4], tt2[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = cb + 8 * j + 2 * t;
        const float2 tv = *reinterpret_cast<const float2*>(tau + c);
        const float2 wv = *reinterpret_cast<const float2*>(w2 + c);
        cb2[j] = *reinterpret_cast<const float2*>(cbias + c);
        tw2[j] = make_float2(tv.x * wv.x, tv.y * wv.y);
        tt2[j] = make_float2(tv.x * tv.x, tv.y * tv.y);
      }
      if (threadIdx.x == 0) N2NMN_STAMP(2, 25);
#pragma unroll 1   // (two row tiles in flight were measured slower: registers)
      for (int mt = pass; mt < pass_end; ++mt) {
        if (threadIdx.x == 0 && mt - pass < 3) N2NMN_STAMP(2, 17 + (mt - pass));
        const int p0 = mt * 16 + g, p1 = p0 + 8;
        const int y0 = (int)(((uint32_t)p0 * inv_w) >> 20), y1 = (int)(((uint32_t)p1 * inv_w) >> 20);
        const int base0 = (p0 < HW) ? y0 * PW + (p0 - y0 * Ww) : 0;
        const int base1 = (p1 < HW) ? y1 * PW + (p1 - y1 * Ww) : 0;
        float acc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
          uint32_t af[4];
          af[0] = __float_as_uint(pad[base0 + toff[ks][0]]);
          af[1] = __float_as_uint(pad[base1 + toff[ks][0]]);
          af[2] = __float_as_uint(pad[base0 + toff[ks][1]]);
          af[3] = __float_as_uint(pad[base1 + toff[ks][1]]);
#pragma unroll
          for (int j = 0; j < 4; ++j) mma_tf32_m16n8k8(acc[j], af, bf[ks][j]);
        }
        // rows g and g+8: num = Σ_c (m+b)·τ·w2, den = Σ_c ((m+b)·τ)²
        float n0 = 0.f, d0 = 0.f, n1 = 0.f, d1 = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = acc[j][0] + cb2[j].x, b = acc[j][1] + cb2[j].y;
          const float c2 = acc[j][2] + cb2[j].x, d = acc[j][3] + cb2[j].y;
          n0 = fmaf(a, tw2[j].x, n0); n0 = fmaf(b, tw2[j].y, n0);
          d0 = fmaf(a * a, tt2[j].x, d0); d0 = fmaf(b * b, tt2[j].y, d0);
          n1 = fmaf(c2, tw2[j].x, n1); n1 = fmaf(d, tw2[j].y, n1);
          d1 = fmaf(c2 * c2, tt2[j].x, d1); d1 = fmaf(d * d, tt2[j].y, d1);
        }
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
          n0 += __shfl_xor_sync(0xffffffffu, n0, o); d0 += __shfl_xor_sync(0xffffffffu, d0, o);
          n1 += __shfl_xor_sync(0xffffffffu, n1, o); d1 += __shfl_xor_sync(0xffffffffu, d1, o);
        }
        if (t == 0) {
          float* r0 = scratch + ((warp * kStencilRowsPerPass) + (mt - pass) * 16 + g) * 2;
          float* r1 = r0 + 16;
          if (first_cb) { r0[0] = n0; r0[1] = d0; r1[0] = n1; r1[1] = d1; }
          else { r0[0] += n0; r0[1] += d0; r1[0] += n1; r1[1] += d1; }
        }
      }
      first_cb = false;
    }
    if (threadIdx.x == 0) N2NMN_STAMP(2, 31);
    __syncthreads();
    for (int r = threadIdx.x; r < (pass_end - pass) * 16; r += blockDim.x) {
      const int p = pass * 16 + r;
      if (p < HW) {
        float n = 0.f, d = 0.f;
        for (int w = 0; w < nwarps; ++w) {
          if (w * 32 < Mp) {   // warps beyond the channel range wrote nothing
            n += scratch[(w * kStencilRowsPerPass + r) * 2];
            d += scratch[(w * kStencilRowsPerPass + r) * 2 + 1];
          }
        }
        ob[p] = n * rsqrtf(fmaxf(d, kEps)) + b2;
      }
    }
    __syncthreads();
  }
}

// kDirect: evaluation schedules with pooled_direct — Describe / SameProperty roots only write their
// pooled feature vectors, the head kernel (head_kernel.cuh) finishes them.
template <int KS, bool kDirect>
__global__ void __launch_bounds__(kNodeThreads, 2)
tree_kernel(const NodeCtx c, const NodeRec* __restrict__ nodes, const int32_t* __restrict__ q_ptr,
            int csize, int stack_slots, int flags) {
  const bool write_arena = (flags & kTreeWriteArena) != 0;
  const bool fp32_stencil = (flags & kTreeFp32Stencil) != 0;
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
      // the filter bank: as is for the fp32 stencil, rounded to TF32 on the way in for the tensor
      // path (this runs under the predecessors' tail, before griddepcontrol.wait)
      if (fp32_stencil) {
        stage_async(s.k, md.conv_k, KS * KS * Mp);
      } else {
        const float4* src = reinterpret_cast<const float4*>(md.conv_k);
        float4* dst = reinterpret_cast<float4*>(s.k);
        for (int j = threadIdx.x; j < KS * KS * Mp / 4; j += blockDim.x) {
          const float4 v = __ldg(src + j);
          dst[j] = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
        }
      }
      for (int ch = threadIdx.x; ch < Mp; ch += blockDim.x) {
        s_tw2[ch] = (ch < M) ? md.elt_w[ES_TRANSFORM][ch] : 0.f;
        s_tcb[ch] = (ch < M) ? md.conv_b[ch] : 0.f;
      }
    }
    if (co.rank == 0 && s.head != nullptr &&
        !(kDirect && (nodes[end - 1].op == OP_DESCRIBE || nodes[end - 1].op == OP_SAME_PROPERTY))) {
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
      for (int i = threadIdx.x; i < md.C; i += blockDim.x) score_row(c, q)[i] = 0.f;
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
        for (int j = threadIdx.x; j < 2 * PH * PW; j += blockDim.x) s.pad[j] = 0.f;
        for (int ch = threadIdx.x; ch < Mp; ch += blockDim.x) {
          s.v0[ch] = (ch < M) ? tau[ch] : 0.f;
          s.v1[ch] = s_tw2[ch];
          s.v2[ch] = s_tcb[ch];
        }
        __syncthreads();
        for (int p = threadIdx.x; p < HW; p += blockDim.x) {
          const int y = p / Ww, x = p - y * Ww;
          s.pad[(y + R) * PW + x + R] = fp32_stencil ? in0[p] : round_tf32(in0[p]);
        }
        cp_async_commit_wait_all();   // the filter bank (prologue)
        __syncthreads();

        if (threadIdx.x == 0) N2NMN_STAMP(2, 21);
        const float b2 = md.elt_b[ES_TRANSFORM][0];
        float* ob = s.outbuf + (exch & 1) * L.HWp;
        if (!fp32_stencil) {
          transform_stencil_mma<KS>(co, s.pad, s.k, s.v0, s.v1, s.v2, b2, s.scratch, ob, Hh, Ww, Mp);
        } else {
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
        }
        if (threadIdx.x == 0) N2NMN_STAMP(2, 22);
        co.sync();
        if (threadIdx.x == 0) N2NMN_STAMP(2, 23);
        if (!fp32_stencil) {
          for (int p = threadIdx.x; p < HW; p += blockDim.x)
            out[p] = co.peer(ob, stencil_owner(p, HW, co.size))[p];
        } else {
          gather_pixels(co, ob, out, HW, Ww, kTransformPB);
        }
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
        // four pixels per step: their map rows are independent L2 reads, issued together (one pixel
        // at a time the warp pays an L2 round trip per pixel: ~19 of them with one CTA per question)
        constexpr int PX = 4;
        for (int pb = gwarp; pb < HW; pb += gwarps * PX) {
          float num[PX], den[PX];
#pragma unroll
          for (int u = 0; u < PX; ++u) { num[u] = 0.f; den[u] = 0.f; }
          for (int qd = lane; qd < (Mp >> 2); qd += 32) {
            float4 m[PX];
#pragma unroll
            for (int u = 0; u < PX; ++u) {
              const int p = min(pb + u * gwarps, HW - 1);
              m[u] = __ldg(reinterpret_cast<const float4*>(mimg + (size_t)p * Mp) + qd);
            }
            const float4 a = reinterpret_cast<const float4*>(s.v1)[qd];
            const float4 d = reinterpret_cast<const float4*>(s.v2)[qd];
#pragma unroll
            for (int u = 0; u < PX; ++u) {
              num[u] = fmaf(m[u].x, a.x, num[u]); num[u] = fmaf(m[u].y, a.y, num[u]);
              num[u] = fmaf(m[u].z, a.z, num[u]); num[u] = fmaf(m[u].w, a.w, num[u]);
              den[u] = fmaf(m[u].x * m[u].x, d.x, den[u]); den[u] = fmaf(m[u].y * m[u].y, d.y, den[u]);
              den[u] = fmaf(m[u].z * m[u].z, d.z, den[u]); den[u] = fmaf(m[u].w * m[u].w, d.w, den[u]);
            }
          }
#pragma unroll
          for (int u = 0; u < PX; ++u) {
            const float n = warp_sum(num[u]), dd = warp_sum(den[u]);
            const int p = pb + u * gwarps;
            if (lane == 0 && p < HW) ob[p] = n * rsqrtf(fmaxf(dd, kEps)) + b2;
          }
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
        if (kDirect) {
          // evaluation: hand the (already softmaxed) attention weights to the pooling kernel
          // (head_kernel.cuh: f = Σ_p s_p·X_b[p,:], then fc_att / l2norm / fc_eltwise batched over
          // many root nodes); this question is done
          if (co.rank == 0) {
            float* w0 = c.pool_att + (size_t)nd.aux * L.HWp;
            for (int p = threadIdx.x; p < HW; p += blockDim.x) w0[p] = s.a0[p];
            if (two) {
              float* w1 = c.pool_att + (size_t)nd.aux2 * L.HWp;
              for (int p = threadIdx.x; p < HW; p += blockDim.x) w1[p] = s.a1[p];
            }
          }
          break;
        }
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
                   score_row(c, nd.out), s.scratch);
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
                 score_row(c, nd.out), s.scratch);
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
