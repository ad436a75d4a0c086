// the attentional seq2seq layout generator of the reference,
// AttentionSeq2Seq (models_clevr/nmn3_netgen_att.py:46-322; the VQA and SHAPES copies are the
// same code), inference configuration: encoder = embedding + multi-layer LSTM under
// dynamic_rnn (:73-120), decoder = raw_rnn loop with tanh attention over the encoder outputs,
// token scores, the Assembler's validity masks, greedy decoding or teacher forcing (:122-322).
// Produces on the device what nmn3_model.py consumes: predicted_tokens, token_probs (whose logs
// sum to log_seq_prob), neg_entropy, word_vecs and the attention maps.
//
// Data layout (fp32, time-major like the reference's tensors):
//   table_enc [V_txt][4L]    = embedding_mat · W_x of encoder layer 0 (the embedding lookup and
//   table_dec [V_nmn+1][4L]    the layer-0 input product fold into ONE row gather per token;
//                              row V_nmn of table_dec is go_embedding)
//   w_cell[side][l] [(in + L)][4L], b_cell [4L]: the BasicLSTMCell matrix (rows of the layer
//       input, then rows of the recurrent state — TF's order) with the gate columns REGROUPED per
//       8 units (regroup_gates_kernel) so that the cell update happens in the GEMM epilogue;
//   h[l] double buffered [2][N][L] (every CTA reads all of h_prev while others write h_next),
//   c[l] [N][L] in place; enc_out / enc_ht [T][N][L]; atts [T_dec][T_enc][N].
// Kernels: lstm_step (one launch per layer per time step: [x, h_prev]·W on mma.sync m16n8k8 with
// error-compensated TF32 = fp32 parity, 32 or 64 questions x 8 units per CTA, K streamed through
// a cp.async ring, LSTM cell in the epilogue), s2s_gemm (same tile engine: h-transform, attention
// query, table precompute), dec_attn (one CTA per question and step: attention, context vector,
// token scores, validity mask, argmax / forcing, probabilities, entropy, stack-state update),
// word_vecs. The whole call is a chain of (T_enc + layers - 1) + layers·T_dec + 2·T_dec + 3
// dependent launches on the caller's stream (the encoder layers run as a wavefront), linked by
// programmatic dependent launch: each kernel requests what does not depend on its predecessor
// (weights, tables) before griddepcontrol.wait.
// At N = 64 an LSTM launch costs the issue time of its mma.sync instructions (three passes for fp32
// parity): the next step for this row is a persistent tcgen05 kernel with the
// weights resident in shared memory.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/n2nmn_b200.h"
#include "mma_tile.cuh"

namespace n2nmn {
int fail_with(int code, const std::string& msg);   // capi.cu: sets n2nmn_last_error()
}
using namespace n2nmn;

namespace {

#define S2S_TRY(expr)                                                                        \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return fail_with(N2NMN_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));  \
  } while (0)

constexpr int kMaxLayers = 4;
constexpr int kAttnThreads = 512;
constexpr int kMaxVocabNmn = 64;    // token scores / masks live in one warp's reach
constexpr int kMaxTEnc = 128;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// Gate columns regrouped per 8 units: dst column (u/8)*32 + g*8 + u%8 = src column g*L + u
// (gate g = i, j, f, o of unit u). A CTA of the step kernel owns 32 such columns = all four gates
// of 8 units, and inside it n-tile g of the mma holds gate g: one thread's accumulators across
// the four n-tiles are the four gates of its two units.
__global__ void regroup_gates_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                     int rows, int L) {
  const int n = rows * 4 * L;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = i / (4 * L), c = i - r * 4 * L;
    const int blk = c >> 5, g = (c >> 3) & 3, u = blk * 8 + (c & 7);
    dst[i] = src[(size_t)r * 4 * L + g * L + u];
  }
}

__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows,
                                 int cols) {   // dst[c][r] = src[r][c]
  const int n = rows * cols;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = i / rows, r = i - c * rows;
    dst[i] = src[(size_t)r * cols + c];
  }
}

// out[r][c] = Σ_k A[r][k] B[k][c] + bias[c];  grid = (ceil(C/32), ceil(R/(16 WM)))
template <int WM, bool kExact>
__global__ void __launch_bounds__(kMmaThreads)
s2s_gemm_kernel(GemmOperands p, const float* __restrict__ bias, float* __restrict__ out, int ldo) {
  pdl_trigger();
  extern __shared__ __align__(16) float mma_smem[];
  const int row0 = blockIdx.y * 16 * WM, c0 = blockIdx.x * kMmaCols;
  float acc[4][4];
  if (!mma_tile<WM, kExact>(mma_smem, p, row0, c0, acc, [] {})) return;
  const int lane = threadIdx.x & 31, wm = (threadIdx.x >> 5) % WM, g = lane >> 2, tig = lane & 3;
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int r = row0 + wm * 16 + g + 8 * hh;
    if (r >= p.R) continue;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = c0 + nt * 8 + 2 * tig + j;
        if (c < p.C) out[(size_t)r * ldo + c] = acc[nt][hh * 2 + j] + (bias ? bias[c] : 0.f);
      }
  }
}

struct LstmStep {
  const float* x;        // [N][L] output of the layer below at this step, or nullptr (layer 0)
  const float* h_prev;   // [N][L]
  const float* w;        // [(L +) L][4L] regrouped: rows of x (layer >= 1) then rows of h_prev
  const float* table;    // layer 0: [V][4L] regrouped input products, else nullptr
  const int32_t* tok;    // layer 0: token of question n at this step
  const float* bias;     // [4L] regrouped
  float* c;              // [N][L] in place
  float* h_out;          // [N][L]
  float* out_seq;        // encoder top layer: encoder_outputs[t] (zero past the end) or nullptr
  const int32_t* seq_len;   // encoder: [N]; nullptr in the decoder (every row live)
  int t, N, L;
};

// BasicLSTMCell(forget_bias=1) step (gate order i, j, f, o) with dynamic_rnn's masking: past the
// sequence end the state is carried through and the output is zero (nmn3_netgen_att.py:95-99).
// grid = (4L/32, ceil(N/(16 WM))): one CTA = 8 units x 64 (WM = 4) or 32 (WM = 2) questions; the
// narrow variant is used while it is what it takes to put a CTA on most SMs (N <= 64 at L = 512).
// The encoder runs as a WAVEFRONT: layer l of time step t and layer l-1 of step t+1 only depend on
// the previous launch, so one launch carries up to kMaxLayers steps (blockIdx.z = layer; a slot with
// N == 0 is idle) and the encoder takes T + layers - 1 launches instead of T * layers. With the
// 3-stage ring two CTAs share an SM, which keeps as many bytes in flight as the 5-stage ring of
// the single-step launches.
struct LstmWave { LstmStep s[kMaxLayers]; };
template <int WM, bool kExact, int ST>
__global__ void __launch_bounds__(kMmaThreads) lstm_step_kernel(const LstmWave wave) {
  pdl_trigger();
  const LstmStep& p = wave.s[blockIdx.z];
  if (p.N == 0) return;
  extern __shared__ __align__(16) float mma_smem[];
  const int row0 = blockIdx.y * 16 * WM, c0 = blockIdx.x * kMmaCols;
  const int L = p.L, C = 4 * L;
  GemmOperands op;
  if (p.x != nullptr) { op.a0 = p.x; op.k0 = L; op.lda0 = L; op.a1 = p.h_prev; op.k1 = L; op.lda1 = L; }
  else { op.a0 = p.h_prev; op.k0 = L; op.lda0 = L; op.a1 = nullptr; op.k1 = 0; op.lda1 = 0; }
  op.R = p.N; op.B = p.w; op.ldb = C; op.C = C;
  float acc[4][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wm = warp % WM, g = lane >> 2,
            tig = lane & 3;
  // epilogue inputs, requested as soon as the previous step is complete (they ride under the GEMM)
  float gt[2][4][2], c_prev[2][2], h_keep[2][2];
  bool live[2];
  auto prefetch = [&] {
    if (warp / WM != 0) return;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int n = row0 + wm * 16 + g + 8 * hh;
#pragma unroll
      for (int gate = 0; gate < 4; ++gate) {
        const float2 b2 = *reinterpret_cast<const float2*>(p.bias + c0 + gate * 8 + 2 * tig);
        gt[hh][gate][0] = b2.x; gt[hh][gate][1] = b2.y;
      }
      live[hh] = false;
      if (n >= p.N) continue;
      if (p.table != nullptr) {
        const float* row = p.table + (size_t)p.tok[n] * C + c0 + 2 * tig;
#pragma unroll
        for (int gate = 0; gate < 4; ++gate) {
          const float2 e = *reinterpret_cast<const float2*>(row + gate * 8);
          gt[hh][gate][0] += e.x; gt[hh][gate][1] += e.y;
        }
      }
      live[hh] = p.seq_len == nullptr || p.t < p.seq_len[n];
      const size_t idx = (size_t)n * L + (c0 >> 2) + 2 * tig;
      const float2 cp = *reinterpret_cast<const float2*>(p.c + idx);
      const float2 hp = *reinterpret_cast<const float2*>(p.h_prev + idx);
      c_prev[hh][0] = cp.x; c_prev[hh][1] = cp.y;
      h_keep[hh][0] = hp.x; h_keep[hh][1] = hp.y;
    }
  };
  if (!mma_tile<WM, kExact, ST>(mma_smem, op, row0, c0, acc, prefetch)) return;
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int n = row0 + wm * 16 + g + 8 * hh;
    if (n >= p.N) continue;
    float c_new[2], h_new[2], o_new[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float gi = gt[hh][0][j] + acc[0][hh * 2 + j], gj = gt[hh][1][j] + acc[1][hh * 2 + j];
      const float gf = gt[hh][2][j] + acc[2][hh * 2 + j], go = gt[hh][3][j] + acc[3][hh * 2 + j];
      const float c2 = c_prev[hh][j] * sigmoidf_(gf + 1.0f) + sigmoidf_(gi) * tanhf(gj);
      const float h2 = tanhf(c2) * sigmoidf_(go);
      c_new[j] = live[hh] ? c2 : c_prev[hh][j];
      h_new[j] = live[hh] ? h2 : h_keep[hh][j];
      o_new[j] = live[hh] ? h2 : 0.f;
    }
    const size_t idx = (size_t)n * L + (c0 >> 2) + 2 * tig;
    *reinterpret_cast<float2*>(p.c + idx) = make_float2(c_new[0], c_new[1]);
    *reinterpret_cast<float2*>(p.h_out + idx) = make_float2(h_new[0], h_new[1]);
    if (p.out_seq != nullptr)
      *reinterpret_cast<float2*>(p.out_seq + idx) = make_float2(o_new[0], o_new[1]);
  }
}

struct AttnStep {
  const float* q;         // [N][L] h_top · W_a + b_a
  const float* h_top;     // [N][L]
  const float* enc_ht;    // [T][N][L]
  const float* enc_out;   // [T][N][L]
  const float* v;         // [L]
  const float* wy_t;      // [V][2L] token_prediction weights, transposed
  const float* by;        // [V]
  const int32_t* seq_len; // [N]
  const int32_t* P;       // [V][3]
  const int32_t* W;       // [3][V][4]
  const int32_t* b;       // [V][4]
  int32_t* X;             // [N][3] decoding state (nmn3_netgen_att.py:288-293)
  const int32_t* gt;      // [N] this step's ground-truth tokens or nullptr
  const float* u;         // [N] this step's uniform numbers (decoder_sampling) or nullptr
  int32_t* tokens;        // [N] this step's predicted tokens (row t of predicted_tokens)
  int32_t* cur_tok;       // [N] input token of the next step
  float* probs;           // [N] row t of token_probs
  float* neg_entropy;     // [N] accumulated
  float* atts;            // [T][N] this step's attention
  int T, N, L, V;
};

// One CTA per question: nmn3_netgen_att.py:205-293 for one decoding step. Everything that does
// not depend on the previous kernel (W_y^T, v, b_y and the Assembler tables) is staged in shared
// memory BEFORE griddepcontrol.wait, i.e. under the tail of the kernels before it. The step is a
// chain of short phases, each bound by the latency of its global loads, so every phase keeps as
// many independent 16-byte loads in flight per thread as registers allow.
__host__ __device__ inline size_t attn_smem_floats(int L, int T, int V) {
  const int part = L > 4 * kAttnThreads ? L : 4 * kAttnThreads;
  return (size_t)V * 2 * L + 4 * L + part + ((T + 3) & ~3) + 2 * ((V + 3) & ~3) + 3 * V + 12 * V +
         4 * V + 8;
}
__global__ void __launch_bounds__(kAttnThreads) dec_attn_kernel(AttnStep p) {
  pdl_trigger();
  extern __shared__ __align__(16) float sm[];
  const int n = blockIdx.x, L = p.L, T = p.T, V = p.V;
  const int part = L > 4 * kAttnThreads ? L : 4 * kAttnThreads;
  float* s_wy = sm;                    // [V][2L]
  float* s_x = s_wy + (size_t)V * 2 * L;   // [2L] = [h_top, d2]
  float* s_q = s_x + 2 * L;            // [L]
  float* s_v = s_q + L;                // [L]
  float* s_part = s_v + L;             // [G][L] partial context vectors
  float* s_att = s_part + part;        // [T]
  float* s_sc = s_att + ((T + 3) & ~3);   // [V] scores
  float* s_by = s_sc + ((V + 3) & ~3);    // [V]
  int32_t* s_P = reinterpret_cast<int32_t*>(s_by + ((V + 3) & ~3));   // [V][3]
  int32_t* s_W = s_P + 3 * V;          // [3][V][4]
  int32_t* s_b = s_W + 12 * V;         // [V][4]
  int32_t* s_valid = s_b + 4 * V;      // [2]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = kAttnThreads / 32;
  for (int i = tid; i < V * 2 * L / 4; i += kAttnThreads) tp_cp16(s_wy + 4 * i, p.wy_t + 4 * i);
  tp_commit();
  for (int d = tid; d < L; d += kAttnThreads) s_v[d] = p.v[d];
  for (int i = tid; i < V; i += kAttnThreads) s_by[i] = p.by[i];
  for (int i = tid; i < 3 * V; i += kAttnThreads) s_P[i] = p.P[i];
  for (int i = tid; i < 12 * V; i += kAttnThreads) s_W[i] = p.W[i];
  for (int i = tid; i < 4 * V; i += kAttnThreads) s_b[i] = p.b[i];
  pdl_wait();
  for (int d = tid; d < L; d += kAttnThreads) {
    s_x[d] = p.h_top[(size_t)n * L + d];
    s_q[d] = p.q[(size_t)n * L + d];
  }
  __syncthreads();
  const int ncol = L >> 2;
  const size_t tstride = (size_t)p.N * L;
  // att_raw[te] = Σ_d tanh(q + enc_ht[te]) v   (:208-212): a warp takes 4 time steps at once
  for (int tb = warp * 4; tb < T; tb += nw * 4) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    const float* ht = p.enc_ht + (size_t)tb * tstride + (size_t)n * L;
#pragma unroll 2
    for (int c4 = lane; c4 < ncol; c4 += 32) {
      const float4 q4 = reinterpret_cast<const float4*>(s_q)[c4];
      const float4 v4 = reinterpret_cast<const float4*>(s_v)[c4];
      float4 h4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        h4[i] = tb + i < T ? __ldg(reinterpret_cast<const float4*>(ht + i * tstride) + c4)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        s[i] += tanhf(q4.x + h4[i].x) * v4.x + tanhf(q4.y + h4[i].y) * v4.y +
                tanhf(q4.z + h4[i].z) * v4.z + tanhf(q4.w + h4[i].w) * v4.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int o = 16; o; o >>= 1) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
      if (lane == 0 && tb + i < T) s_att[tb + i] = s[i];
    }
  }
  __syncthreads();
  // softmax over ALL time steps, then mask by the sequence length and renormalise (:213-216)
  if (warp == 0) {
    float m = -INFINITY;
    for (int te = lane; te < T; te += 32) m = fmaxf(m, s_att[te]);
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int te = lane; te < T; te += 32) { const float e = expf(s_att[te] - m); s_att[te] = e; s += e; }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const int len = p.seq_len[n];
    float s2 = 0.f;
    for (int te = lane; te < T; te += 32) {
      const float a = te < len ? s_att[te] / s : 0.f;
      s_att[te] = a; s2 += a;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    for (int te = lane; te < T; te += 32) {
      const float a = s_att[te] / s2;
      s_att[te] = a;
      p.atts[(size_t)te * p.N + n] = a;
    }
  } else if (warp == 1) {
    // validity of every token from the decoding state (:8-11); all ones under forcing (:230-233)
    const int32_t x0 = p.X[n * 3], x1 = p.X[n * 3 + 1], x2 = p.X[n * 3 + 2];
    uint32_t lo = 0, hi = 0;
    for (int base = 0; base < V; base += 32) {
      const int vv = base + lane;
      bool ok = vv < V;
      if (ok && p.gt == nullptr) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int32_t lhs = x0 * s_W[(0 * V + vv) * 4 + c] + x1 * s_W[(1 * V + vv) * 4 + c] +
                              x2 * s_W[(2 * V + vv) * 4 + c];
          ok = ok && (lhs - s_b[vv * 4 + c] >= 0);
        }
      }
      const uint32_t bal = __ballot_sync(0xffffffffu, ok);
      if (base == 0) lo = bal; else hi = bal;
    }
    if (lane == 0) { s_valid[0] = (int32_t)lo; s_valid[1] = (int32_t)hi; }
  }
  __syncthreads();
  // d2 = Σ_te att[te] encoder_outputs[te]   (:218): G groups of threads split the time steps of
  // each 4-channel column, partial sums meet in shared memory
  {
    const int G = ncol >= kAttnThreads ? 1 : kAttnThreads / ncol;
    for (int item = tid; item < ncol * G; item += kAttnThreads) {
      const int c4 = item % ncol, gi = item / ncol;
      const float4* eo = reinterpret_cast<const float4*>(p.enc_out + (size_t)n * L) + c4;
      float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
      for (int te = gi; te < T; te += G) {
        const float4 e = __ldg(eo + (size_t)te * (tstride >> 2));
        const float a = s_att[te];
        acc4.x += a * e.x; acc4.y += a * e.y; acc4.z += a * e.z; acc4.w += a * e.w;
      }
      reinterpret_cast<float4*>(s_part + (size_t)gi * L)[c4] = acc4;
    }
    __syncthreads();
    for (int d = tid; d < L; d += kAttnThreads) {
      float s = 0.f;
      for (int gi = 0; gi < G; ++gi) s += s_part[(size_t)gi * L + d];
      s_x[L + d] = s;
    }
  }
  tp_wait<0>();
  __syncthreads();
  // token_scores = [h_top, d2] · W_y + b_y   (:221-223)
  for (int vv = warp; vv < V; vv += nw) {
    const float4* wr = reinterpret_cast<const float4*>(s_wy + (size_t)vv * 2 * L);
    float s = 0.f;
    for (int k4 = lane; k4 < 2 * ncol; k4 += 32) {
      const float4 x4 = reinterpret_cast<const float4*>(s_x)[k4], w4 = wr[k4];
      s += x4.x * w4.x + x4.y * w4.y + x4.z * w4.z + x4.w * w4.w;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_sc[vv] = s + s_by[vv];
  }
  __syncthreads();
  if (warp == 0) {   // lane handles tokens lane and lane + 32 (V <= 64)
    const uint32_t vlo = (uint32_t)s_valid[0], vhi = (uint32_t)s_valid[1];
    const int v0 = lane, v1 = lane + 32;
    const bool in0 = v0 < V, in1 = v1 < V;
    const bool ok0 = in0 && ((vlo >> lane) & 1), ok1 = in1 && ((vhi >> lane) & 1);
    const float sc0 = in0 ? s_sc[v0] : -INFINITY, sc1 = in1 ? s_sc[v1] : -INFINITY;
    float mx = fmaxf(sc0, sc1);
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    // greedy token: invalid scores are replaced by (global min - 1) before the argmax (:259-261),
    // i.e. the first best VALID token, or token 0 if none is valid
    float best = ok0 ? sc0 : -INFINITY;
    int pred = ok0 ? v0 : 0x7fffffff;
    if (ok1 && sc1 > best) { best = sc1; pred = v1; }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int op = __shfl_xor_sync(0xffffffffu, pred, o);
      if (ob > best || (ob == best && op < pred)) { best = ob; pred = op; }
    }
    if (pred == 0x7fffffff) pred = 0;
    if (p.u != nullptr) {
      // decoder_sampling (:234-256): one draw from softmax(scores - 50·invalid) by inverse CDF
      // over the tokens in vocabulary order — the first token whose cumulative probability
      // exceeds u — kept when it is valid, else the greedy token above
      const float z0 = in0 ? sc0 - (ok0 ? 0.f : 50.f) : -INFINITY;
      const float z1 = in1 ? sc1 - (ok1 ? 0.f : 50.f) : -INFINITY;
      float zm = fmaxf(z0, z1);
#pragma unroll
      for (int o = 16; o; o >>= 1) zm = fmaxf(zm, __shfl_xor_sync(0xffffffffu, zm, o));
      float c0 = in0 ? expf(z0 - zm) : 0.f, c1 = in1 ? expf(z1 - zm) : 0.f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {   // inclusive scans over the lanes
        const float t0 = __shfl_up_sync(0xffffffffu, c0, o), t1 = __shfl_up_sync(0xffffffffu, c1, o);
        if (lane >= o) { c0 += t0; c1 += t1; }
      }
      c1 += __shfl_sync(0xffffffffu, c0, 31);
      const float thr = p.u[n] * __shfl_sync(0xffffffffu, c1, 31);
      const uint32_t m0 = __ballot_sync(0xffffffffu, in0 && c0 > thr);
      const uint32_t m1 = __ballot_sync(0xffffffffu, in1 && c1 > thr);
      const int samp = m0 ? __ffs(m0) - 1 : (m1 ? 31 + __ffs(m1) : V - 1);
      if (((samp < 32 ? vlo >> samp : vhi >> (samp - 32)) & 1u) != 0u) pred = samp;
    }
    if (p.gt != nullptr) pred = p.gt[n];   // :264-266
    const float e0 = in0 ? expf(sc0 - mx) : 0.f, e1 = in1 ? expf(sc1 - mx) : 0.f;
    float se = e0 + e1;
#pragma unroll
    for (int o = 16; o; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    const float a0 = ok0 ? e0 / se : 0.f, a1 = ok1 ? e1 / se : 0.f;      // :270
    float sv = a0 + a1;
#pragma unroll
    for (int o = 16; o; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
    const float p0 = a0 / sv, p1 = a1 / sv;                               // :272
    float ent = 0.f;
    if (in0) ent += p0 * logf(fmaxf(1e-5f, p0 + (ok0 ? 0.f : 1.f)));      // :283-285
    if (in1) ent += p1 * logf(fmaxf(1e-5f, p1 + (ok1 ? 0.f : 1.f)));
#pragma unroll
    for (int o = 16; o; o >>= 1) ent += __shfl_xor_sync(0xffffffffu, ent, o);
    const float pp = __shfl_sync(0xffffffffu, pred < 32 ? p0 : p1, pred & 31);
    if (lane == 0) {
      p.probs[n] = pp;                                                    // :281
      p.neg_entropy[n] += ent;
      p.X[n * 3] += s_P[pred * 3];                                        // :288-289
      p.X[n * 3 + 1] += s_P[pred * 3 + 1];
      p.X[n * 3 + 2] += s_P[pred * 3 + 2];
      p.tokens[n] = pred;
      p.cur_tok[n] = pred;
    }
  }
}

// word_vecs[td][n][:] = Σ_te atts[td][te][n] · embedding_mat[input_seq[te][n]]   (:312)
// grid = (N, T_dec)
__global__ void word_vecs_kernel(const float* __restrict__ atts, const int32_t* __restrict__ seq,
                                 const float* __restrict__ emb, float* __restrict__ out, int T,
                                 int N, int E) {
  const int n = blockIdx.x, td = blockIdx.y;
  extern __shared__ float s_a[];   // [T] weights then [T] token ids (as int)
  int32_t* s_tok = reinterpret_cast<int32_t*>(s_a + T);
  for (int te = threadIdx.x; te < T; te += blockDim.x) {
    s_a[te] = atts[((size_t)td * T + te) * N + n];
    s_tok[te] = seq[(size_t)te * N + n];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float s = 0.f;
    for (int te = 0; te < T; ++te) s += s_a[te] * emb[(size_t)s_tok[te] * E + e];
    out[((size_t)td * N + n) * E + e] = s;
  }
}

__global__ void init_state_kernel(int32_t* X, int32_t* cur_tok, float* neg_entropy, int N,
                                  int T_dec, int go_row) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  X[n * 3] = 0; X[n * 3 + 1] = 0; X[n * 3 + 2] = T_dec;   // :293
  cur_tok[n] = go_row;
  neg_entropy[n] = 0.f;
}

struct S2SVar {
  std::string name;
  std::vector<int64_t> shape;
  float* dev;       // raw copy as given (TF layout)
  size_t count;
  bool loaded;
};

}  // namespace

struct n2nmn_seq2seq {
  n2nmn_seq2seq_config cfg;
  std::vector<S2SVar> vars;
  bool dirty = true, tables_set = false;
  const float* sample_u = nullptr;   // [T_decoder][N] uniforms of the following forward calls
  // derived weights
  float *table_enc = nullptr, *table_dec = nullptr, *dec_rows = nullptr;   // dec_rows = [emb; go]
  float* w_cell[2][kMaxLayers] = {};   // interleaved full matrices [(in+L)][4L]
  float* b_cell[2][kMaxLayers] = {};
  float* wy_t = nullptr;
  // state / workspaces
  float* h[kMaxLayers][2] = {};
  float* c[kMaxLayers] = {};
  float *enc_out = nullptr, *enc_ht = nullptr, *q = nullptr, *atts = nullptr;
  int32_t *X = nullptr, *cur_tok = nullptr, *P = nullptr, *W = nullptr, *b = nullptr;
  int64_t launches = 0;
  int var(const std::string& n) const {
    for (size_t i = 0; i < vars.size(); ++i) if (vars[i].name == n) return (int)i;
    return -1;
  }
  const float* v(const std::string& n) const { return vars[var(n)].dev; }
};

namespace {

template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                       cudaStream_t st, Args... args) {
  cudaLaunchConfig_t lc;
  std::memset(&lc, 0, sizeof(lc));
  lc.gridDim = grid; lc.blockDim = block; lc.dynamicSmemBytes = smem; lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = 1;
  return cudaLaunchKernelEx(&lc, kernel, KArgs(args)...);
}

// 32-row tiles while 64-row tiles would leave SMs without a CTA
bool narrow_tiles(int col_blocks, int R) { return R > 16 && col_blocks * ((R + 63) / 64) < 148; }

int launch_gemm(n2nmn_seq2seq* s, cudaStream_t st, const float* A, int lda, int R, int K,
                const float* B, int ldb, int C, const float* bias, float* out, int ldo,
                bool force_exact = false) {
  GemmOperands op;
  op.a0 = A; op.k0 = K; op.lda0 = lda; op.a1 = nullptr; op.k1 = 0; op.lda1 = 0;
  op.R = R; op.B = B; op.ldb = ldb; op.C = C;
  const int cb = (C + kMmaCols - 1) / kMmaCols;
  const bool exact = !(s->cfg.flags & N2NMN_SEQ2SEQ_FLAG_TF32) || force_exact;
  const dim3 gn(cb, (R + 31) / 32), gw(cb, (R + 63) / 64), blk(kMmaThreads);
  if (narrow_tiles(cb, R)) {
    if (exact) S2S_TRY(launch_pdl(s2s_gemm_kernel<2, true>, gn, blk, mma_smem_bytes(2), st, op, bias, out, ldo));
    else S2S_TRY(launch_pdl(s2s_gemm_kernel<2, false>, gn, blk, mma_smem_bytes(2), st, op, bias, out, ldo));
  } else {
    if (exact) S2S_TRY(launch_pdl(s2s_gemm_kernel<4, true>, gw, blk, mma_smem_bytes(4), st, op, bias, out, ldo));
    else S2S_TRY(launch_pdl(s2s_gemm_kernel<4, false>, gw, blk, mma_smem_bytes(4), st, op, bias, out, ldo));
  }
  ++s->launches;
  return N2NMN_OK;
}

std::string cell_prefix(int side, int l) {
  return std::string(side == 0 ? "encoder" : "decoder") + "/lstm/multi_rnn_cell/cell_" +
         std::to_string(l) + "/basic_lstm_cell/";
}

// Re-derive the packed weights after a set_weight (once; on the caller's stream).
int prepare(n2nmn_seq2seq* s, cudaStream_t st) {
  const auto& g = s->cfg;
  const int L = g.lstm_dim, C = 4 * L;
  for (auto& v : s->vars)
    if (!v.loaded) return fail_with(N2NMN_ERR_STATE, "seq2seq weight not set: " + v.name);
  for (int side = 0; side < 2; ++side) {
    for (int l = 0; l < g.num_layers; ++l) {
      const int in = l == 0 ? (side == 0 ? g.embed_dim_txt : g.embed_dim_nmn) : L;
      regroup_gates_kernel<<<148, 256, 0, st>>>(s->v(cell_prefix(side, l) + "weights"),
                                                s->w_cell[side][l], in + L, L);
      regroup_gates_kernel<<<8, 256, 0, st>>>(s->v(cell_prefix(side, l) + "biases"),
                                              s->b_cell[side][l], 1, L);
      s->launches += 2;
    }
  }
  S2S_TRY(cudaMemcpyAsync(s->dec_rows, s->v("decoder/embedding_mat"),
                          sizeof(float) * g.num_vocab_nmn * g.embed_dim_nmn,
                          cudaMemcpyDeviceToDevice, st));
  S2S_TRY(cudaMemcpyAsync(s->dec_rows + (size_t)g.num_vocab_nmn * g.embed_dim_nmn,
                          s->v("decoder/go_embedding"), sizeof(float) * g.embed_dim_nmn,
                          cudaMemcpyDeviceToDevice, st));
  int rc = launch_gemm(s, st, s->v("encoder/embedding_mat"), g.embed_dim_txt, g.num_vocab_txt,
                       g.embed_dim_txt, s->w_cell[0][0], C, C, nullptr, s->table_enc, C, true);
  if (rc) return rc;
  rc = launch_gemm(s, st, s->dec_rows, g.embed_dim_nmn, g.num_vocab_nmn + 1, g.embed_dim_nmn,
                   s->w_cell[1][0], C, C, nullptr, s->table_dec, C, true);
  if (rc) return rc;
  transpose_kernel<<<64, 256, 0, st>>>(s->v("decoder/token_prediction/weights"), s->wy_t, 2 * L,
                                       g.num_vocab_nmn);
  ++s->launches;
  S2S_TRY(cudaGetLastError());
  s->dirty = false;
  return N2NMN_OK;
}

template <class T>
cudaError_t dmalloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)); }

}  // namespace

extern "C" {

int n2nmn_seq2seq_create(const n2nmn_seq2seq_config* cfg, n2nmn_seq2seq** out) {
  if (!cfg || !out) return fail_with(N2NMN_ERR_ARG, "null argument");
  if (cfg->abi_version != N2NMN_ABI_VERSION) return fail_with(N2NMN_ERR_ARG, "ABI version mismatch");
  const int L = cfg->lstm_dim;
  if (cfg->num_vocab_txt <= 0 || cfg->embed_dim_txt <= 0 || cfg->embed_dim_nmn <= 0 ||
      cfg->embed_dim_txt % 4 != 0 || cfg->embed_dim_nmn % 4 != 0 ||
      cfg->num_vocab_nmn <= 0 || cfg->num_vocab_nmn > kMaxVocabNmn || L <= 0 || L % 16 != 0 ||
      cfg->num_layers <= 0 || cfg->num_layers > kMaxLayers || cfg->T_encoder <= 0 ||
      cfg->T_encoder > kMaxTEnc || cfg->T_decoder <= 0 || cfg->max_batch <= 0)
    return fail_with(N2NMN_ERR_ARG,
                     "bad seq2seq config (lstm_dim must be a multiple of 16, embed dims of 4, num_vocab_nmn <= 64, "
                     "num_layers <= 4, T_encoder <= 128)");
  S2S_TRY(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  S2S_TRY(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10)
    return fail_with(N2NMN_ERR_DEVICE, std::string("n2nmn_b200 needs an sm_100 GPU, found sm_") +
                                           std::to_string(prop.major) + std::to_string(prop.minor));
  auto* s = new n2nmn_seq2seq;
  s->cfg = *cfg;
  const int C = 4 * L, N = cfg->max_batch, Vt = cfg->num_vocab_txt, Vn = cfg->num_vocab_nmn;
  const int Et = cfg->embed_dim_txt, En = cfg->embed_dim_nmn;
  auto add = [&](const std::string& name, std::vector<int64_t> shape) {
    size_t cnt = 1;
    for (auto d : shape) cnt *= (size_t)d;
    s->vars.push_back({name, shape, nullptr, cnt, false});
  };
  // variables of `encoder_decoder/` in creation order (nmn3_netgen_att.py:73-240)
  add("encoder/embedding_mat", {Vt, Et});
  for (int l = 0; l < cfg->num_layers; ++l) {
    add(cell_prefix(0, l) + "weights", {(l == 0 ? Et : L) + L, C});
    add(cell_prefix(0, l) + "biases", {C});
  }
  add("encoder/encoder_h_transform/weights", {L, L});
  add("encoder/encoder_h_transform/biases", {L});
  add("decoder/embedding_mat", {Vn, En});
  add("decoder/go_embedding", {1, En});
  add("decoder/att_prediction/v", {L});
  add("decoder/att_prediction/weights", {L, L});
  add("decoder/att_prediction/biases", {L});
  add("decoder/token_prediction/weights", {2 * L, Vn});
  add("decoder/token_prediction/biases", {Vn});
  for (int l = 0; l < cfg->num_layers; ++l) {
    add(cell_prefix(1, l) + "weights", {(l == 0 ? En : L) + L, C});
    add(cell_prefix(1, l) + "biases", {C});
  }
  for (auto& v : s->vars) S2S_TRY(dmalloc(&v.dev, v.count));
  S2S_TRY(dmalloc(&s->table_enc, (size_t)Vt * C));
  S2S_TRY(dmalloc(&s->table_dec, (size_t)(Vn + 1) * C));
  S2S_TRY(dmalloc(&s->dec_rows, (size_t)(Vn + 1) * En));
  S2S_TRY(dmalloc(&s->wy_t, (size_t)Vn * 2 * L));
  for (int side = 0; side < 2; ++side)
    for (int l = 0; l < cfg->num_layers; ++l) {
      const int in = l == 0 ? (side == 0 ? Et : En) : L;
      S2S_TRY(dmalloc(&s->w_cell[side][l], (size_t)(in + L) * C));
      S2S_TRY(dmalloc(&s->b_cell[side][l], (size_t)C));
    }
  for (int l = 0; l < cfg->num_layers; ++l) {
    S2S_TRY(dmalloc(&s->h[l][0], (size_t)N * L));
    S2S_TRY(dmalloc(&s->h[l][1], (size_t)N * L));
    S2S_TRY(dmalloc(&s->c[l], (size_t)N * L));
  }
  const size_t TNL = (size_t)cfg->T_encoder * N * L;
  S2S_TRY(dmalloc(&s->enc_out, TNL));
  S2S_TRY(dmalloc(&s->enc_ht, TNL));
  S2S_TRY(dmalloc(&s->q, (size_t)N * L));
  S2S_TRY(dmalloc(&s->atts, (size_t)cfg->T_decoder * cfg->T_encoder * N));
  S2S_TRY(dmalloc(&s->X, (size_t)N * 3));
  S2S_TRY(dmalloc(&s->cur_tok, (size_t)N));
  S2S_TRY(dmalloc(&s->P, (size_t)Vn * 3));
  S2S_TRY(dmalloc(&s->W, (size_t)3 * Vn * 4));
  S2S_TRY(dmalloc(&s->b, (size_t)Vn * 4));
  auto opt_in = [](auto kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  };
  S2S_TRY(opt_in(s2s_gemm_kernel<2, true>, mma_smem_bytes(2)));
  S2S_TRY(opt_in(s2s_gemm_kernel<2, false>, mma_smem_bytes(2)));
  S2S_TRY(opt_in(s2s_gemm_kernel<4, true>, mma_smem_bytes(4)));
  S2S_TRY(opt_in(s2s_gemm_kernel<4, false>, mma_smem_bytes(4)));
  S2S_TRY(opt_in(lstm_step_kernel<2, true, 5>, mma_smem_bytes(2, 5)));
  S2S_TRY(opt_in(lstm_step_kernel<2, false, 5>, mma_smem_bytes(2, 5)));
  S2S_TRY(opt_in(lstm_step_kernel<2, true, 3>, mma_smem_bytes(2, 3)));
  S2S_TRY(opt_in(lstm_step_kernel<2, false, 3>, mma_smem_bytes(2, 3)));
  S2S_TRY(opt_in(lstm_step_kernel<4, true, 3>, mma_smem_bytes(4, 3)));
  S2S_TRY(opt_in(lstm_step_kernel<4, false, 3>, mma_smem_bytes(4, 3)));
  S2S_TRY(cudaFuncSetAttribute(lstm_step_kernel<2, true, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  S2S_TRY(cudaFuncSetAttribute(lstm_step_kernel<2, false, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  const size_t attn_bytes = attn_smem_floats(L, cfg->T_encoder, Vn) * sizeof(float);
  if (attn_bytes > 200 * 1024)
    return fail_with(N2NMN_ERR_ARG, "num_vocab_nmn * lstm_dim too large for the decoder step kernel");
  S2S_TRY(cudaFuncSetAttribute(dec_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)attn_bytes));
  *out = s;
  return N2NMN_OK;
}

int n2nmn_seq2seq_destroy(n2nmn_seq2seq* s) {
  if (!s) return N2NMN_OK;
  for (auto& v : s->vars) cudaFree(v.dev);
  cudaFree(s->table_enc); cudaFree(s->table_dec); cudaFree(s->dec_rows); cudaFree(s->wy_t);
  for (int side = 0; side < 2; ++side)
    for (int l = 0; l < kMaxLayers; ++l) { cudaFree(s->w_cell[side][l]); cudaFree(s->b_cell[side][l]); }
  for (int l = 0; l < kMaxLayers; ++l) { cudaFree(s->h[l][0]); cudaFree(s->h[l][1]); cudaFree(s->c[l]); }
  cudaFree(s->enc_out); cudaFree(s->enc_ht); cudaFree(s->q); cudaFree(s->atts);
  cudaFree(s->X); cudaFree(s->cur_tok); cudaFree(s->P); cudaFree(s->W); cudaFree(s->b);
  delete s;
  return N2NMN_OK;
}

int n2nmn_seq2seq_num_variables(const n2nmn_seq2seq* s) { return s ? (int)s->vars.size() : 0; }

int n2nmn_seq2seq_variable_info(const n2nmn_seq2seq* s, int index, const char** name,
                                int64_t shape[4], int* ndim) {
  if (!s || index < 0 || index >= (int)s->vars.size()) return fail_with(N2NMN_ERR_ARG, "bad variable index");
  const auto& v = s->vars[index];
  if (name) *name = v.name.c_str();
  if (ndim) *ndim = (int)v.shape.size();
  if (shape) for (size_t i = 0; i < v.shape.size() && i < 4; ++i) shape[i] = v.shape[i];
  return N2NMN_OK;
}

int n2nmn_seq2seq_set_weight(n2nmn_seq2seq* s, const char* name, const float* src_dev,
                             const int64_t* shape, int ndim, void* stream) {
  if (!s || !name || !src_dev || !shape) return fail_with(N2NMN_ERR_ARG, "null argument");
  const int i = s->var(name);
  if (i < 0) return fail_with(N2NMN_ERR_ARG, std::string("unknown seq2seq variable: ") + name);
  auto& v = s->vars[i];
  bool same = ndim == (int)v.shape.size();
  for (int d = 0; same && d < ndim; ++d) same = shape[d] == v.shape[d];
  if (!same) return fail_with(N2NMN_ERR_ARG, std::string("shape mismatch for ") + name);
  S2S_TRY(cudaMemcpyAsync(v.dev, src_dev, v.count * sizeof(float), cudaMemcpyDeviceToDevice,
                          (cudaStream_t)stream));
  v.loaded = true;
  s->dirty = true;
  return N2NMN_OK;
}

int n2nmn_seq2seq_set_assembler(n2nmn_seq2seq* s, const int32_t* P, const int32_t* W,
                                const int32_t* b, void* stream) {
  if (!s || !P || !W || !b) return fail_with(N2NMN_ERR_ARG, "null argument");
  const int Vn = s->cfg.num_vocab_nmn;
  auto st = (cudaStream_t)stream;
  S2S_TRY(cudaMemcpyAsync(s->P, P, sizeof(int32_t) * Vn * 3, cudaMemcpyHostToDevice, st));
  S2S_TRY(cudaMemcpyAsync(s->W, W, sizeof(int32_t) * 3 * Vn * 4, cudaMemcpyHostToDevice, st));
  S2S_TRY(cudaMemcpyAsync(s->b, b, sizeof(int32_t) * Vn * 4, cudaMemcpyHostToDevice, st));
  S2S_TRY(cudaStreamSynchronize(st));   // the host arrays may be temporaries
  s->tables_set = true;
  return N2NMN_OK;
}

int n2nmn_seq2seq_forward(n2nmn_seq2seq* s, const int32_t* input_seq_dev,
                          const int32_t* seq_len_dev, int T_enc, int N,
                          const int32_t* gt_layout_dev, int32_t* tokens_dev,
                          float* token_probs_dev, float* neg_entropy_dev, float* word_vecs_dev,
                          float* atts_dev, void* stream) {
  if (!s || !input_seq_dev || !seq_len_dev || !tokens_dev || !token_probs_dev ||
      !neg_entropy_dev || !word_vecs_dev)
    return fail_with(N2NMN_ERR_ARG, "null argument");
  const auto& g = s->cfg;
  if (T_enc <= 0 || T_enc > g.T_encoder || N <= 0 || N > g.max_batch)
    return fail_with(N2NMN_ERR_CAPACITY, "T_enc / N exceed what the seq2seq was created for");
  if (!s->tables_set) return fail_with(N2NMN_ERR_STATE, "assembler tables (P, W, b) not set");
  auto st = (cudaStream_t)stream;
  if (s->dirty) {
    const int rc = prepare(s, st);
    if (rc) return rc;
  }
  const int L = g.lstm_dim, C = 4 * L, NL = g.num_layers, Vn = g.num_vocab_nmn;
  const int Et = g.embed_dim_txt, En = g.embed_dim_nmn, T_dec = g.T_decoder;
  float* atts = atts_dev ? atts_dev : s->atts;
  for (int l = 0; l < NL; ++l) {
    S2S_TRY(cudaMemsetAsync(s->h[l][0], 0, sizeof(float) * N * L, st));
    S2S_TRY(cudaMemsetAsync(s->c[l], 0, sizeof(float) * N * L, st));
  }
  init_state_kernel<<<(N + 127) / 128, 128, 0, st>>>(s->X, s->cur_tok, neg_entropy_dev, N, T_dec, Vn);
  ++s->launches;
  const bool narrow = narrow_tiles(C / kMmaCols, N);
  const dim3 grid(C / kMmaCols, narrow ? (N + 31) / 32 : (N + 63) / 64);
  int cur = 0;   // h[l][cur] holds every layer's h_{t-1}
  bool ok = true;
  const bool exact = !(g.flags & N2NMN_SEQ2SEQ_FLAG_TF32);
  // one LSTM cell evaluation: layer l of `side` at step t; `par` = parity of the h buffer that
  // holds the layer's previous state (its own and the layer below's run in lock step)
  auto cell = [&](int side, int l, int t, int par, const int32_t* tok, const int32_t* seq_len,
                  float* out_seq) {
    const int in = l == 0 ? (side == 0 ? Et : En) : L;
    LstmStep p;
    p.x = l == 0 ? nullptr : s->h[l - 1][par ^ 1];
    p.h_prev = s->h[l][par];
    p.w = s->w_cell[side][l] + (l == 0 ? (size_t)in * C : 0);
    p.table = l == 0 ? (side == 0 ? s->table_enc : s->table_dec) : nullptr;
    p.tok = tok;
    p.bias = s->b_cell[side][l];
    p.c = s->c[l];
    p.h_out = s->h[l][par ^ 1];
    p.out_seq = l == NL - 1 ? out_seq : nullptr;
    p.seq_len = seq_len;
    p.t = t; p.N = N; p.L = L;
    return p;
  };
  auto launch_wave = [&](const LstmWave& w, int nz, bool shared_sm) {
    const dim3 g3(grid.x, grid.y, nz), blk(kMmaThreads);
    cudaError_t le;
    if (!narrow) {
      le = exact ? launch_pdl(lstm_step_kernel<4, true, 3>, g3, blk, mma_smem_bytes(4, 3), st, w)
                 : launch_pdl(lstm_step_kernel<4, false, 3>, g3, blk, mma_smem_bytes(4, 3), st, w);
    } else if (shared_sm) {
      le = exact ? launch_pdl(lstm_step_kernel<2, true, 3>, g3, blk, mma_smem_bytes(2, 3), st, w)
                 : launch_pdl(lstm_step_kernel<2, false, 3>, g3, blk, mma_smem_bytes(2, 3), st, w);
    } else {
      le = exact ? launch_pdl(lstm_step_kernel<2, true, 5>, g3, blk, mma_smem_bytes(2, 5), st, w)
                 : launch_pdl(lstm_step_kernel<2, false, 5>, g3, blk, mma_smem_bytes(2, 5), st, w);
    }
    if (le != cudaSuccess) ok = false;
    ++s->launches;
  };
  // encoder, dynamic_rnn (:95-99): tick k runs layer l at step t = k - l
  for (int k = 0; k < T_enc + NL - 1; ++k) {
    LstmWave w;
    std::memset(&w, 0, sizeof(w));
    for (int l = 0; l < NL; ++l) {
      const int t = k - l;
      if (t < 0 || t >= T_enc) continue;    // idle slot (N = 0)
      w.s[l] = cell(0, l, t, t & 1, input_seq_dev + (size_t)t * N, seq_len_dev,
                    s->enc_out + (size_t)t * N * L);
    }
    launch_wave(w, NL, NL > 1);
  }
  cur = T_enc & 1;
  // decoder step: the layers of one step depend on each other, one launch each
  auto step = [&](int t, const int32_t* tok) {
    for (int l = 0; l < NL; ++l) {
      LstmWave w;
      std::memset(&w, 0, sizeof(w));
      w.s[0] = cell(1, l, t, cur, tok, nullptr, nullptr);
      launch_wave(w, 1, false);
    }
    cur ^= 1;
  };
  S2S_TRY(cudaGetLastError());
  int rc = launch_gemm(s, st, s->enc_out, L, T_enc * N, L, s->v("encoder/encoder_h_transform/weights"),
                       L, L, s->v("encoder/encoder_h_transform/biases"), s->enc_ht, L);   // :104-108
  if (rc) return rc;
  const size_t attn_smem = attn_smem_floats(L, T_enc, Vn) * sizeof(float);
  for (int t = 0; t < T_dec; ++t) {   // raw_rnn loop (:199-305)
    step(t, s->cur_tok);
    const float* h_top = s->h[NL - 1][cur];
    rc = launch_gemm(s, st, h_top, L, N, L, s->v("decoder/att_prediction/weights"), L, L,
                     s->v("decoder/att_prediction/biases"), s->q, L);
    if (rc) return rc;
    AttnStep a;
    a.q = s->q; a.h_top = h_top; a.enc_ht = s->enc_ht; a.enc_out = s->enc_out;
    a.v = s->v("decoder/att_prediction/v");
    a.wy_t = s->wy_t; a.by = s->v("decoder/token_prediction/biases");
    a.seq_len = seq_len_dev; a.P = s->P; a.W = s->W; a.b = s->b; a.X = s->X;
    a.gt = gt_layout_dev ? gt_layout_dev + (size_t)t * N : nullptr;
    a.u = s->sample_u ? s->sample_u + (size_t)t * N : nullptr;
    a.tokens = tokens_dev + (size_t)t * N;
    a.cur_tok = s->cur_tok;
    a.probs = token_probs_dev + (size_t)t * N;
    a.neg_entropy = neg_entropy_dev;
    a.atts = atts + (size_t)t * T_enc * N;
    a.T = T_enc; a.N = N; a.L = L; a.V = Vn;
    if (launch_pdl(dec_attn_kernel, dim3(N), dim3(kAttnThreads), attn_smem, st, a) != cudaSuccess)
      ok = false;
    ++s->launches;
  }
  word_vecs_kernel<<<dim3(N, T_dec), 128, sizeof(float) * 2 * T_enc, st>>>(
      atts, input_seq_dev, s->v("encoder/embedding_mat"), word_vecs_dev, T_enc, N, Et);
  ++s->launches;
  S2S_TRY(cudaGetLastError());
  if (!ok) return fail_with(N2NMN_ERR_CUDA, "seq2seq kernel launch failed");
  return N2NMN_OK;
}

int n2nmn_seq2seq_set_sampling(n2nmn_seq2seq* s, const float* uniforms_dev) {
  if (!s) return fail_with(N2NMN_ERR_ARG, "null argument");
  s->sample_u = uniforms_dev;
  return N2NMN_OK;
}

int64_t n2nmn_seq2seq_launch_count(const n2nmn_seq2seq* s) { return s ? s->launches : 0; }

}  // extern "C"
