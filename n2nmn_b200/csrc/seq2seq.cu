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
//   w_in[l] [L][4L] (l >= 1), w_rec[l] [L][4L], bias[l] [4L] : the BasicLSTMCell matrix split by
//       input, with gate columns INTERLEAVED (column 4u+g = gate g of unit u, g = i, j, f, o) so
//       that one thread's 4 adjacent accumulators are the 4 gates of one unit and the cell update
//       happens in the GEMM epilogue;
//   h[l] double buffered [2][N][L] (every CTA reads all of h_prev while others write h_next),
//   c[l] [N][L] in place; enc_out / enc_ht [T][N][L]; atts [T_dec][T_enc][N].
// Kernels: lstm_step (one per layer per time step; 64x64-tile fp32 GEMM over [h_below, h_prev]
// with the whole K in shared memory, text_proj.cuh), s2s_gemm (h-transform, attention query,
// table precompute), dec_attn (one CTA per question and step: attention, context vector, token
// scores, validity mask, argmax / forcing, probabilities, entropy, stack state update),
// word_vecs. A step is latency bound (N <= 64 rows): ~32 CTAs; the whole call is a chain of
// 2·L_layers·(T_enc + T_dec) + 2·T_dec + 2 launches captured in order on the caller's stream.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/n2nmn_b200.h"
#include "tile_gemm.cuh"

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
constexpr int kAttnThreads = 256;
constexpr int kMaxVocabNmn = 64;    // token scores / masks live in one warp's reach
constexpr int kMaxTEnc = 128;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// dst[r][4u+g] = src[r][g*L+u]
__global__ void interleave_gates_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                        int rows, int L) {
  const int n = rows * 4 * L;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = i / (4 * L), c = i - r * 4 * L;
    const int u = c >> 2, g = c & 3;
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

// out[r][c] = Σ_k A[r][k] B[k][c] + bias[c];  grid = (ceil(C/64), ceil(R/64))
__global__ void __launch_bounds__(kTileThreads)
s2s_gemm_kernel(const float* __restrict__ A, int lda, int R, int K, const float* __restrict__ B,
                int ldb, int C, const float* __restrict__ bias, float* __restrict__ out, int ldo) {
  extern __shared__ __align__(16) float tile_smem[];
  const int row0 = blockIdx.y * kTileRows, c0 = blockIdx.x * kTextCols;
  auto a_row = [&](int r) -> const float* {
    return row0 + r < R ? A + (size_t)(row0 + r) * lda : nullptr;
  };
  float acc[2][4];
  tile_gemm_64x64(tile_smem, a_row, K, B, ldb, c0, C, 1 << 30, acc);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = row0 + 2 * ty + i;
    if (r >= R) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + 4 * tx + j;
      if (c < C) out[(size_t)r * ldo + c] = acc[i][j] + (bias ? bias[c] : 0.f);
    }
  }
}

struct LstmStep {
  const float* x;        // [N][L] output of the layer below at this step, or nullptr (layer 0)
  const float* w_in;     // [L][4L] interleaved (layer >= 1)
  const float* h_prev;   // [N][L]
  const float* w_rec;    // [L][4L] interleaved
  const float* table;    // layer 0: [V][4L] interleaved input products, else nullptr
  const int32_t* tok;    // layer 0: token of question n at this step
  const float* bias;     // [4L] interleaved
  float* c;              // [N][L] in place
  float* h_out;          // [N][L]
  float* out_seq;        // encoder top layer: encoder_outputs[t] (zero past the end) or nullptr
  const int32_t* seq_len;   // encoder: [N]; nullptr in the decoder (every row live)
  int t, N, L;
};

// BasicLSTMCell(forget_bias=1) step (gate order i, j, f, o) with dynamic_rnn's masking: past the
// sequence end the state is carried through and the output is zero (nmn3_netgen_att.py:95-99).
// grid = (4L/64, ceil(N/64))
__global__ void __launch_bounds__(kTileThreads) lstm_step_kernel(LstmStep p) {
  extern __shared__ __align__(16) float tile_smem[];
  const int row0 = blockIdx.y * kTileRows, c0 = blockIdx.x * kTextCols;
  const int L = p.L, C = 4 * L;
  float acc[2][4];
  auto h_row = [&](int r) -> const float* {
    return row0 + r < p.N ? p.h_prev + (size_t)(row0 + r) * L : nullptr;
  };
  tile_gemm_64x64(tile_smem, h_row, L, p.w_rec, C, c0, C, 1 << 30, acc);
  if (p.x != nullptr) {
    auto x_row = [&](int r) -> const float* {
      return row0 + r < p.N ? p.x + (size_t)(row0 + r) * L : nullptr;
    };
    tile_gemm_64x64(tile_smem, x_row, L, p.w_in, C, c0, C, 1 << 30, acc, false);
  }
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int col = c0 + 4 * tx, u = col >> 2;
  if (col >= C) return;
  const float4 b4 = *reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int n = row0 + 2 * ty + i;
    if (n >= p.N) continue;
    float g[4] = {acc[i][0] + b4.x, acc[i][1] + b4.y, acc[i][2] + b4.z, acc[i][3] + b4.w};
    if (p.table != nullptr) {
      const float4 e = *reinterpret_cast<const float4*>(p.table + (size_t)p.tok[n] * C + col);
      g[0] += e.x; g[1] += e.y; g[2] += e.z; g[3] += e.w;
    }
    const size_t idx = (size_t)n * L + u;
    const float c_prev = p.c[idx];
    const float c2 = c_prev * sigmoidf_(g[2] + 1.0f) + sigmoidf_(g[0]) * tanhf(g[1]);
    const float h2 = tanhf(c2) * sigmoidf_(g[3]);
    const bool live = p.seq_len == nullptr || p.t < p.seq_len[n];
    p.c[idx] = live ? c2 : c_prev;
    p.h_out[idx] = live ? h2 : p.h_prev[idx];
    if (p.out_seq != nullptr) p.out_seq[idx] = live ? h2 : 0.f;
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
  int32_t* tokens;        // [N] this step's predicted tokens (row t of predicted_tokens)
  int32_t* cur_tok;       // [N] input token of the next step
  float* probs;           // [N] row t of token_probs
  float* neg_entropy;     // [N] accumulated
  float* atts;            // [T][N] this step's attention
  int T, N, L, V;
};

// One CTA per question: nmn3_netgen_att.py:205-293 for one decoding step.
__global__ void __launch_bounds__(kAttnThreads) dec_attn_kernel(AttnStep p) {
  extern __shared__ __align__(16) float sm[];
  const int n = blockIdx.x, L = p.L, T = p.T, V = p.V;
  float* s_x = sm;                 // [2L] = [h_top, d2]
  float* s_q = sm + 2 * L;         // [L]
  float* s_v = s_q + L;            // [L]
  float* s_att = s_v + L;          // [T]
  float* s_sc = s_att + T;         // [V]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = kAttnThreads / 32;
  for (int d = threadIdx.x; d < L; d += kAttnThreads) {
    s_x[d] = p.h_top[(size_t)n * L + d];
    s_q[d] = p.q[(size_t)n * L + d];
    s_v[d] = p.v[d];
  }
  __syncthreads();
  // att_raw[te] = Σ_d tanh(q + enc_ht[te]) v   (:208-212)
  for (int te = warp; te < T; te += nw) {
    const float* ht = p.enc_ht + ((size_t)te * p.N + n) * L;
    float s = 0.f;
    for (int d = lane; d < L; d += 32) s += tanhf(s_q[d] + ht[d]) * s_v[d];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_att[te] = s;
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
  }
  __syncthreads();
  // d2 = Σ_te att[te] encoder_outputs[te]   (:218)
  for (int d = threadIdx.x; d < L; d += kAttnThreads) {
    float s = 0.f;
    for (int te = 0; te < T; ++te) s += s_att[te] * p.enc_out[((size_t)te * p.N + n) * L + d];
    s_x[L + d] = s;
  }
  __syncthreads();
  // token_scores = [h_top, d2] · W_y + b_y   (:221-223)
  for (int vv = warp; vv < V; vv += nw) {
    const float* wr = p.wy_t + (size_t)vv * 2 * L;
    float s = 0.f;
    for (int k = lane; k < 2 * L; k += 32) s += s_x[k] * wr[k];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_sc[vv] = s + p.by[vv];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int32_t x0 = p.X[n * 3], x1 = p.X[n * 3 + 1], x2 = p.X[n * 3 + 2];
    uint64_t valid = 0;
    float mx = -INFINITY;
    for (int vv = 0; vv < V; ++vv) {
      bool ok = true;
      if (p.gt == nullptr) {   // _get_valid_tokens (:8-11); all ones under teacher forcing (:230-233)
        for (int c = 0; c < 4; ++c) {
          const int32_t lhs = x0 * p.W[(0 * V + vv) * 4 + c] + x1 * p.W[(1 * V + vv) * 4 + c] +
                              x2 * p.W[(2 * V + vv) * 4 + c];
          ok = ok && (lhs - p.b[vv * 4 + c] >= 0);
        }
      }
      if (ok) valid |= 1ull << vv;
      mx = fmaxf(mx, s_sc[vv]);
    }
    // greedy token: invalid scores are replaced by (global min - 1) before the argmax (:259-261),
    // i.e. the first best VALID token, or token 0 if none is valid
    int pred = 0;
    float best = -INFINITY;
    for (int vv = 0; vv < V; ++vv)
      if (((valid >> vv) & 1) && s_sc[vv] > best) { best = s_sc[vv]; pred = vv; }
    if (p.gt != nullptr) pred = p.gt[n];   // :264-266
    float se = 0.f;
    for (int vv = 0; vv < V; ++vv) { const float e = expf(s_sc[vv] - mx); s_sc[vv] = e; se += e; }
    float sv = 0.f;
    for (int vv = 0; vv < V; ++vv) {
      const float a = ((valid >> vv) & 1) ? s_sc[vv] / se : 0.f;   // :270
      s_sc[vv] = a; sv += a;
    }
    float ent = 0.f;
    for (int vv = 0; vv < V; ++vv) {
      const float a = s_sc[vv] / sv;                               // :272
      s_sc[vv] = a;
      const float inv = ((valid >> vv) & 1) ? 0.f : 1.f;
      ent += a * logf(fmaxf(1e-5f, a + inv));                      // :283-285
    }
    p.probs[n] = s_sc[pred];                                        // :281
    p.neg_entropy[n] += ent;
    p.X[n * 3] = x0 + p.P[pred * 3];                                // :288-289
    p.X[n * 3 + 1] = x1 + p.P[pred * 3 + 1];
    p.X[n * 3 + 2] = x2 + p.P[pred * 3 + 2];
    p.tokens[n] = pred;
    p.cur_tok[n] = pred;
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

size_t gemm_smem(int K) { return (size_t)tile_smem_floats(K) * 4 + kTileRows * sizeof(void*); }

int launch_gemm(n2nmn_seq2seq* s, cudaStream_t st, const float* A, int lda, int R, int K,
                const float* B, int ldb, int C, const float* bias, float* out, int ldo) {
  dim3 grid((C + kTextCols - 1) / kTextCols, (R + kTileRows - 1) / kTileRows);
  s2s_gemm_kernel<<<grid, kTileThreads, gemm_smem(K), st>>>(A, lda, R, K, B, ldb, C, bias, out, ldo);
  ++s->launches;
  S2S_TRY(cudaGetLastError());
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
      interleave_gates_kernel<<<148, 256, 0, st>>>(s->v(cell_prefix(side, l) + "weights"),
                                                   s->w_cell[side][l], in + L, L);
      interleave_gates_kernel<<<8, 256, 0, st>>>(s->v(cell_prefix(side, l) + "biases"),
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
                       g.embed_dim_txt, s->w_cell[0][0], C, C, nullptr, s->table_enc, C);
  if (rc) return rc;
  rc = launch_gemm(s, st, s->dec_rows, g.embed_dim_nmn, g.num_vocab_nmn + 1, g.embed_dim_nmn,
                   s->w_cell[1][0], C, C, nullptr, s->table_dec, C);
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
      cfg->num_vocab_nmn <= 0 || cfg->num_vocab_nmn > kMaxVocabNmn || L <= 0 || L % 16 != 0 ||
      cfg->num_layers <= 0 || cfg->num_layers > kMaxLayers || cfg->T_encoder <= 0 ||
      cfg->T_encoder > kMaxTEnc || cfg->T_decoder <= 0 || cfg->max_batch <= 0)
    return fail_with(N2NMN_ERR_ARG,
                     "bad seq2seq config (lstm_dim must be a multiple of 16, num_vocab_nmn <= 64, "
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
  const int maxK = std::max(std::max(Et, En), L);
  S2S_TRY(cudaFuncSetAttribute(s2s_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)gemm_smem(maxK)));
  S2S_TRY(cudaFuncSetAttribute(lstm_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)gemm_smem(L)));
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
  const dim3 grid(C / kTextCols, (N + kTileRows - 1) / kTileRows);
  const size_t smem = gemm_smem(L);
  int cur = 0;   // h[l][cur] holds every layer's h_{t-1}
  auto step = [&](int side, int t, const int32_t* tok, const int32_t* seq_len, float* out_seq) {
    for (int l = 0; l < NL; ++l) {
      const int in = l == 0 ? (side == 0 ? Et : En) : L;
      LstmStep p;
      p.x = l == 0 ? nullptr : s->h[l - 1][cur ^ 1];
      p.w_in = s->w_cell[side][l];
      p.h_prev = s->h[l][cur];
      p.w_rec = s->w_cell[side][l] + (size_t)in * C;
      p.table = l == 0 ? (side == 0 ? s->table_enc : s->table_dec) : nullptr;
      p.tok = tok;
      p.bias = s->b_cell[side][l];
      p.c = s->c[l];
      p.h_out = s->h[l][cur ^ 1];
      p.out_seq = l == NL - 1 ? out_seq : nullptr;
      p.seq_len = seq_len;
      p.t = t; p.N = N; p.L = L;
      lstm_step_kernel<<<grid, kTileThreads, smem, st>>>(p);
      ++s->launches;
    }
    cur ^= 1;
  };
  for (int t = 0; t < T_enc; ++t)   // dynamic_rnn (:95-99)
    step(0, t, input_seq_dev + (size_t)t * N, seq_len_dev, s->enc_out + (size_t)t * N * L);
  S2S_TRY(cudaGetLastError());
  int rc = launch_gemm(s, st, s->enc_out, L, T_enc * N, L, s->v("encoder/encoder_h_transform/weights"),
                       L, L, s->v("encoder/encoder_h_transform/biases"), s->enc_ht, L);   // :104-108
  if (rc) return rc;
  const size_t attn_smem = sizeof(float) * (4 * L + T_enc + Vn);
  for (int t = 0; t < T_dec; ++t) {   // raw_rnn loop (:199-305)
    step(1, t, s->cur_tok, nullptr, nullptr);
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
    a.tokens = tokens_dev + (size_t)t * N;
    a.cur_tok = s->cur_tok;
    a.probs = token_probs_dev + (size_t)t * N;
    a.neg_entropy = neg_entropy_dev;
    a.atts = atts + (size_t)t * T_enc * N;
    a.T = T_enc; a.N = N; a.L = L; a.V = Vn;
    dec_attn_kernel<<<N, kAttnThreads, attn_smem, st>>>(a);
    ++s->launches;
  }
  word_vecs_kernel<<<dim3(N, T_dec), 128, sizeof(float) * 2 * T_enc, st>>>(
      atts, input_seq_dev, s->v("encoder/embedding_mat"), word_vecs_dev, T_enc, N, Et);
  ++s->launches;
  S2S_TRY(cudaGetLastError());
  return N2NMN_OK;
}

int64_t n2nmn_seq2seq_launch_count(const n2nmn_seq2seq* s) { return s ? s->launches : 0; }

}  // extern "C"
