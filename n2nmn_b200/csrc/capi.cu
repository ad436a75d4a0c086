// C ABI of the N2NMN module-network hot path (include/n2nmn_b200.h): context, weight packing,
// input binding, schedule upload and kernel launches. sm_100a only; no CPU fallback.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "../../include/n2nmn_b200.h"
#include "common.cuh"
#include "node_eval.cuh"
#include "prep.cuh"
#include "proj_simt.cuh"
#include "proj_umma.cuh"
#include "schedule.hpp"
#include "text_proj.cuh"
#include "tree_kernel.cuh"
#include "head_kernel.cuh"
#include "backward.cuh"
#include "wgrad_umma.cuh"
#include "head_tail_umma.cuh"

using namespace n2nmn;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CUDA_TRY(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess)                                                              \
      return fail(N2NMN_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));   \
  } while (0)

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

enum VarKind { VK_PLAIN = 0, VK_PROJ_W, VK_PROJ_B, VK_PITCHED };

struct Variable {
  std::string name;
  std::vector<int64_t> shape;
  VarKind kind;
  int set;           // projection set for VK_PROJ_*
  size_t offset;     // float offset of the plain copy inside wbuf
  size_t count;
  const float** slot;   // DevModel pointer to fill (plain copy)
  bool loaded;
};

constexpr int kTableSlots = 4;

struct TableSlot {
  uint8_t* host = nullptr;   // pinned staging
  uint8_t* dev = nullptr;
  uint64_t uid = 0;          // schedule resident here (0 = none)
  cudaEvent_t last_use = nullptr;
};

struct TableOffsets {
  size_t nodes, q_ptr, text_t, text_b, groups, work, img_ptr, node_text, node_out, mslot,
      wave_nodes, bwd_nodes, entry_order, node_entry, entries, text_set_start, labels, head_work, head_list, pool_img,
      total;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

struct n2nmn_sched {
  HostSchedule hs;
  uint64_t uid;
  SchedShape shp;
};

struct n2nmn_ctx {
  n2nmn_config cfg;
  int device = 0, num_sms = 0;
  int HW = 0, Dk = 0, Kp = 0, Mp = 0;
  int G = 1;                   // segments (batches) one set of launches may cover
  int QB = 0;                  // question capacity of one launch: G * max_batch
  SchedShape shp;
  DevModel md;
  std::vector<Variable> vars;
  float* wbuf = nullptr;
  size_t wbuf_floats = 0;
  float* proj_wt[NUM_PROJ_SETS] = {};      // [Mp][Kp], allocated for the sets the family owns
  float* proj_bias[NUM_PROJ_SETS] = {};    // [Mp]
  bool proj_used[NUM_PROJ_SETS] = {};
  int num_store_sets = 0;
  // inputs
  bool bound = false;
  int N = 0, T = 0;
  float* feat_aug = nullptr;   // ctx-owned re-pitched / coordinate-augmented copy
  // workspaces
  TextBufs tb = {nullptr, nullptr, nullptr, nullptr};
  int text_rows_cap = 0;
  float* arena = nullptr;
  int arena_slots = 0;
  float* mbuf = nullptr;
  int mbuf_slots = 0;
  float* pooled = nullptr;     // [2*QB][Kp] pooled feature vectors of Describe / SameProperty roots
  float* pool_att = nullptr;   // [2*QB][HWp] input maps of the answer roots (softmaxed for those)
  float* conv_quad = nullptr;  // [quad_rows][Mp] Transform quadratic-form matrix (common.cuh)
  int head_nn = 16;            // root nodes per head-kernel CTA
  int head_smem_bytes = 0;
  float* scores_tmp = nullptr;
  TableSlot slots[kTableSlots];
  size_t table_cap = 0;
  int next_slot = 0;
  ProjTensorMaps tmaps;
  EncodeTiledFn encode = nullptr;
  int node_smem_bytes = 0;
  int tree_smem_bytes = 0;
  int stack_cap = 0;           // attention-stack slots the tree kernel may use
  int tree_cluster = 0;        // 0 = choose from the batch size; else forced (N2NMN_TREE_CLUSTER)
  bool fp32_stencil = false;   // N2NMN_FP32_STENCIL=1: CUDA-core Transform stencil (A/B timing)
  int text_ctas_per_group = 0; // 0 = one CTA per 64-column block
  int proj_max_ctas = 0;       // 0 = one CTA per SM; else cap of the persistent projection grid
  bool use_pdl = true;         // programmatic dependent launch between the three kernels
  n2nmn_sched module_sched;    // scratch schedule of n2nmn_module_fwd
  n2nmn_sched step_sched;      // scratch schedule of n2nmn_forward_tokens
  // training workspaces (allocated on first use)
  const int32_t* train_labels = nullptr;   // host labels of the step being compiled
  const uint8_t* last_tables = nullptr;    // device tables of the last run_tables
  TableOffsets last_offsets;
  float* dscores = nullptr;
  float* per_sample = nullptr;
  float* dtau = nullptr;
  float* dmap = nullptr;
  float* dstencil = nullptr;
  float* gmap = nullptr;
  float* phi_buf = nullptr;
  WgradMaps wg_maps;                       // tcgen05 weight-gradient kernel (wgrad_umma.cuh)
  bool wg_ok = false;                      // shapes fit it (Mp == 256, Dk % 128 == 0, no re-pitch)
  // many-class answer heads (C > 32): ê rows + score-row addresses of the roots of a launch, and
  // the fc_eltwise matrices with rows pitched to a multiple of 4 floats (cp.async alignment)
  float* ehat = nullptr;
  float** ehat_dst = nullptr;
  float* out_wp[NUM_OUT_SETS] = {};
  int Cp = 0;
  // ... and its tcgen05 form (head_tail_umma.cuh): remainder plane of ê, W_outᵀ planes, maps
  float* ehat_lo = nullptr;
  float* out_wt_hi[NUM_OUT_SETS] = {};
  float* out_wt_lo[NUM_OUT_SETS] = {};
  HeadTailMaps ht_maps[NUM_OUT_SETS];
  int Cpad = 0;   // [max_batch][HW][Mp] scratch of the Transform backward
  int dmap_entries = 0;
  VarSeg* d_segs = nullptr;
  float* d_sumsq = nullptr;
  RepackSeg* d_repack = nullptr;           // fused re-pack tables (n2nmn_load_flat_weights)
  ProjRepack proj_repack;
  int proj_repack_sets = 0;
  float dword_scale = 1.f;                 // n2nmn_set_grad_scale
  GradOffsets go;
  std::vector<int64_t> flat_offset;
  int64_t flat_size = 0;
  // e2e staging
  float* e2e_feat = nullptr;
  float* e2e_wv = nullptr;
  float* e2e_scores = nullptr;
  uint16_t* e2e_feat_f16 = nullptr;   // fp16 staging of host features (host_f16 path)
  // profiling
  bool profiling = false;
  std::vector<cudaEvent_t> ev;
  std::vector<const char*> ev_names;
  int ev_used = 0;
  int64_t launches = 0;
};

namespace {

std::atomic<uint64_t> g_uid{1};

void add_var(n2nmn_ctx* c, const std::string& name, std::vector<int64_t> shape, VarKind kind,
             int set, const float** slot) {
  Variable v;
  v.name = name; v.shape = shape; v.kind = kind; v.set = set; v.slot = slot; v.loaded = false;
  v.count = 1;
  for (int64_t d : shape) v.count *= (size_t)d;
  v.offset = c->wbuf_floats;
  const size_t stored = (kind == VK_PITCHED) ? v.count / (size_t)shape.back() * c->Mp : v.count;
  c->wbuf_floats += (stored + 3) & ~(size_t)3;
  c->vars.push_back(v);
}

// proj_set >= 0: conv_image (repacked for the tensor cores); proj_set == -2: a [rows][M] matrix
// stored with row pitch Mp (fc_text / fc_att) for aligned float4 loads.
void add_layer(n2nmn_ctx* c, const std::string& scope, std::vector<int64_t> wshape,
               const float** wslot, const float** bslot, int proj_set = -1) {
  add_var(c, scope + "/weights", wshape,
          proj_set >= 0 ? VK_PROJ_W : (proj_set == -2 ? VK_PITCHED : VK_PLAIN), proj_set, wslot);
  add_var(c, scope + "/biases", {wshape.back()}, proj_set >= 0 ? VK_PROJ_B : VK_PLAIN, proj_set,
          bslot);
}

// Variable table; order and names match n2nmn_b200/weights.py::variable_shapes.
void build_variables(n2nmn_ctx* c) {
  const n2nmn_config& g = c->cfg;
  DevModel& md = c->md;
  const int64_t D = c->Dk, M = g.map_dim, Dt = g.text_dim, C = g.num_choices, k = g.kernel_size;
  const int64_t HW = c->HW;
  add_layer(c, "FindModule/conv_image", {D, M}, &md.proj_w[PS_FIND], nullptr, PS_FIND);
  add_layer(c, "FindModule/fc_text", {Dt, M}, &md.txt_w[TS_FIND], &md.txt_b[TS_FIND], -2);
  add_layer(c, "FindModule/conv_eltwise", {M, 1}, &md.elt_w[ES_FIND], &md.elt_b[ES_FIND]);
  if (g.family == N2NMN_VQA) {
    add_layer(c, "TransformModule/conv_image", {D, M}, &md.proj_w[PS_FSP_IMG], nullptr, PS_FSP_IMG);
    add_layer(c, "TransformModule/fc_text", {Dt, M}, &md.txt_w[TS_FSP], &md.txt_b[TS_FSP], -2);
    add_layer(c, "TransformModule/fc_att", {D, M}, &md.proj_w[PS_FSP_ATT], nullptr, PS_FSP_ATT);
    add_layer(c, "TransformModule/conv_eltwise", {M, 1}, &md.elt_w[ES_FSP], &md.elt_b[ES_FSP]);
  } else {
    add_layer(c, "TransformModule/conv_maps", {k, k, 1, M}, &md.conv_k, &md.conv_b, -2);
    add_layer(c, "TransformModule/text_fc", {Dt, M}, &md.txt_w[TS_TRANSFORM],
              &md.txt_b[TS_TRANSFORM], -2);
    add_layer(c, "TransformModule/conv_eltwise", {M, 1}, &md.elt_w[ES_TRANSFORM],
              &md.elt_b[ES_TRANSFORM]);
  }
  if (g.family == N2NMN_SHAPES) {
    add_layer(c, "AnswerModule/fc_scores", {3, C}, &md.sc_w[SS_EXIST], &md.sc_b[SS_EXIST]);
    return;
  }
  if (g.family == N2NMN_CLEVR) {
    add_layer(c, "FindSamePropertyModule/conv_image", {D, M}, &md.proj_w[PS_FSP_IMG], nullptr,
              PS_FSP_IMG);
    add_layer(c, "FindSamePropertyModule/fc_text", {Dt, M}, &md.txt_w[TS_FSP], &md.txt_b[TS_FSP], -2);
    add_layer(c, "FindSamePropertyModule/fc_att", {D, M}, &md.proj_w[PS_FSP_ATT], nullptr,
              PS_FSP_ATT);
    add_layer(c, "FindSamePropertyModule/conv_eltwise", {M, 1}, &md.elt_w[ES_FSP],
              &md.elt_b[ES_FSP]);
    add_layer(c, "ExistModule/fc_scores", {3, C}, &md.sc_w[SS_EXIST], &md.sc_b[SS_EXIST]);
    add_layer(c, "CountModule/fc_scores", {HW + 2, C}, &md.sc_w[SS_COUNT], &md.sc_b[SS_COUNT]);
    add_layer(c, "EqualNumModule/fc_scores", {2 * (HW + 2), C}, &md.sc_w[SS_EQUAL],
              &md.sc_b[SS_EQUAL]);
    add_layer(c, "MoreNumModule/fc_scores", {2 * (HW + 2), C}, &md.sc_w[SS_MORE],
              &md.sc_b[SS_MORE]);
    add_layer(c, "LessNumModule/fc_scores", {2 * (HW + 2), C}, &md.sc_w[SS_LESS],
              &md.sc_b[SS_LESS]);
    add_layer(c, "SamePropertyModule/fc_text", {Dt, M}, &md.txt_w[TS_SAMEPROP],
              &md.txt_b[TS_SAMEPROP], -2);
    add_layer(c, "SamePropertyModule/fc_att_0", {D, M}, &md.proj_w[PS_SP_ATT0], nullptr,
              PS_SP_ATT0);
    add_layer(c, "SamePropertyModule/fc_att_1", {D, M}, &md.proj_w[PS_SP_ATT1], nullptr,
              PS_SP_ATT1);
    add_layer(c, "SamePropertyModule/fc_eltwise", {M, C}, &md.out_w[OS_SAMEPROP],
              &md.out_b[OS_SAMEPROP]);
  }
  add_layer(c, "DescribeModule/fc_text", {Dt, M}, &md.txt_w[TS_DESCRIBE], &md.txt_b[TS_DESCRIBE], -2);
  add_layer(c, "DescribeModule/fc_att", {D, M}, &md.proj_w[PS_DESC_ATT], nullptr, PS_DESC_ATT);
  add_layer(c, "DescribeModule/fc_eltwise", {M, C}, &md.out_w[OS_DESCRIBE],
            &md.out_b[OS_DESCRIBE]);
}

int encode_2d(n2nmn_ctx* c, CUtensorMap* map, const float* base, uint64_t inner, uint64_t outer,
              uint64_t pitch_elems, uint32_t box_inner, uint32_t box_outer) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * sizeof(float)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = c->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims,
                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(N2NMN_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(r));
  return 0;
}

// [outer][rows][row_pitch] fp32 viewed as (32 elements, rows, blocks of 32 elements, outer) with box
// (32, box_rows, box_blocks, 1): 128-byte swizzle with 32-byte atoms (the MN-major operand layout
// of 4-byte tensor-core operands), rows beyond `rows` zero-filled.
int encode_mn_blocks(n2nmn_ctx* c, CUtensorMap* map, const float* base, uint64_t cols, uint64_t rows,
                     uint64_t outer, uint64_t row_pitch_elems, uint32_t box_rows,
                     uint32_t box_blocks) {
  cuuint64_t dims[4] = {32, rows, cols / 32, outer};
  cuuint64_t strides[3] = {row_pitch_elems * sizeof(float), 32 * sizeof(float),
                           rows * row_pitch_elems * sizeof(float)};
  cuuint32_t box[4] = {32, box_rows, box_blocks, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = c->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims,
                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(N2NMN_ERR_CUDA, "cuTensorMapEncodeTiled (4d) failed with CUresult " + std::to_string(r));
  return 0;
}

TableOffsets table_offsets(const HostSchedule& S) {
  TableOffsets o;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t at = off; off += (bytes + 15) & ~(size_t)15; return at; };
  o.nodes = take(S.nodes.size() * sizeof(NodeRec));
  o.q_ptr = take(S.q_ptr.size() * 4);
  o.text_t = take(S.text_t.size() * 4);
  o.text_b = take(S.text_b.size() * 4);
  o.groups = take(S.groups.size() * sizeof(TextGroup));
  o.work = take(S.work.size() * sizeof(ProjWork));
  o.img_ptr = take(S.img_ptr.size() * 4);
  o.node_text = take(S.node_text.size() * 4);
  o.node_out = take(S.node_out.size() * 4);
  o.mslot = take(S.mslot.size() * 4);
  o.wave_nodes = take(S.wave_nodes.size() * 4);
  o.bwd_nodes = take(S.bwd_nodes.size() * 4);
  o.entry_order = take(S.entry_order.size() * 4);
  o.node_entry = take(S.node_entry.size() * 4);
  o.entries = take(S.entries.size() * sizeof(BwdEntryHost));
  o.text_set_start = take(S.text_set_start.size() * 4);
  o.labels = take(S.train ? (S.q_ptr.size() - 1) * 4 : 0);
  o.head_work = take(S.head_work.size() * sizeof(HeadWork));
  o.head_list = take(S.head_list.size() * 4);
  o.pool_img = take(S.pool_img.size() * 4);
  o.total = off;
  return o;
}

template <class V>
void put(uint8_t* base, size_t off, const V& v) {
  if (!v.empty()) std::memcpy(base + off, v.data(), v.size() * sizeof(v[0]));
}

void prof_mark(n2nmn_ctx* c, const char* name, cudaStream_t st) {
  if (!c->profiling) return;
  if (c->ev_used >= (int)c->ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    c->ev.push_back(e);
    c->ev_names.push_back(name);
  }
  c->ev_names[c->ev_used] = name;
  cudaEventRecord(c->ev[c->ev_used++], st);
}

// Uploads (if needed) and launches everything for one compiled batch.
int run_tables(n2nmn_ctx* c, n2nmn_sched* sc, float* const* scores_seg, float* arena,
               cudaStream_t st, bool force_wave = false, bool write_arena = false) {
  const bool use_wave = (c->cfg.flags & N2NMN_FLAG_WAVE_EXECUTOR) || force_wave ||
                        sc->hs.max_stack > c->stack_cap;
  if (use_wave && sc->hs.pooled_direct) {
    // the wave executor evaluates Describe / SameProperty from stored fc_att maps: rebuild the
    // derived tables in that form (layouts deeper than the shared-memory stack end up here)
    sc->hs.pooled_direct = false;
    if (int rc = finalize_schedule(sc->shp, sc->hs.N, &sc->hs, false))
      return fail(rc, "finalize_schedule failed");
    sc->uid = g_uid++;
  }
  if (use_wave) build_waves(&sc->hs);
  const HostSchedule& S = sc->hs;
  const TableOffsets o = table_offsets(S);
  if (o.total > c->table_cap)
    return fail(N2NMN_ERR_CAPACITY, "schedule tables exceed the context capacity");
  if ((int)S.text_t.size() > c->text_rows_cap)
    return fail(N2NMN_ERR_CAPACITY, "too many text nodes for this context");
  if (S.num_mslots > c->mbuf_slots)
    return fail(N2NMN_ERR_CAPACITY, "too many stored feature maps for this context");
  // ---- table residency
  TableSlot* slot = nullptr;
  for (int i = 0; i < kTableSlots; ++i)
    if (c->slots[i].uid == sc->uid) slot = &c->slots[i];
  if (!slot) {
    slot = &c->slots[c->next_slot];
    c->next_slot = (c->next_slot + 1) % kTableSlots;
    CUDA_TRY(cudaEventSynchronize(slot->last_use));   // previous tenant no longer in flight
    put(slot->host, o.nodes, S.nodes);
    put(slot->host, o.q_ptr, S.q_ptr);
    put(slot->host, o.text_t, S.text_t);
    put(slot->host, o.text_b, S.text_b);
    put(slot->host, o.groups, S.groups);
    put(slot->host, o.work, S.work);
    put(slot->host, o.img_ptr, S.img_ptr);
    put(slot->host, o.node_text, S.node_text);
    put(slot->host, o.node_out, S.node_out);
    put(slot->host, o.mslot, S.mslot);
    put(slot->host, o.wave_nodes, S.wave_nodes);
    put(slot->host, o.bwd_nodes, S.bwd_nodes);
    put(slot->host, o.entry_order, S.entry_order);
    put(slot->host, o.node_entry, S.node_entry);
    put(slot->host, o.entries, S.entries);
    put(slot->host, o.text_set_start, S.text_set_start);
    put(slot->host, o.head_work, S.head_work);
    put(slot->host, o.head_list, S.head_list);
    put(slot->host, o.pool_img, S.pool_img);
    if (S.train && c->train_labels)
      std::memcpy(slot->host + o.labels, c->train_labels, (S.q_ptr.size() - 1) * 4);
    CUDA_TRY(cudaMemcpyAsync(slot->dev, slot->host, o.total, cudaMemcpyHostToDevice, st));
    slot->uid = sc->uid;
  }
  const uint8_t* d = slot->dev;
  c->last_tables = d;
  c->last_offsets = o;
  const NodeRec* d_nodes = reinterpret_cast<const NodeRec*>(d + o.nodes);
  const int32_t* d_qptr = reinterpret_cast<const int32_t*>(d + o.q_ptr);

  c->ev_used = 0;
  prof_mark(c, "begin", st);
  // every kernel of the step after the first is a programmatic dependent launch of its
  // predecessor (each calls griddepcontrol.wait before it touches the predecessor's output)
  auto pdl_ok = [&]() { return c->use_pdl; };
  // ---- K1 text projections (+ the quadratic-form coefficients of the Transform nodes)
  if (!S.groups.empty()) {
    dim3 grid((unsigned)(c->Mp / kMmaCols), (unsigned)S.groups.size());
    TextSetRows tsr;
    for (int i = 0; i <= NUM_TEXT_SETS; ++i) tsr.start[i] = S.text_set_start[i];
    text_proj_kernel<<<grid, kMmaThreads, mma_smem_bytes(4, kTextStages), st>>>(
        c->md, c->tb, tsr,
        reinterpret_cast<const int32_t*>(d + o.text_t),
        reinterpret_cast<const int32_t*>(d + o.text_b));
    ++c->launches;
    prof_mark(c, "text_proj_kernel", st);
    const int tr0 = S.text_set_start[TS_TRANSFORM];
    const int trn = S.text_set_start[TS_TRANSFORM + 1] - tr0;
    if (trn > 0 && c->conv_quad) {
      cudaLaunchConfig_t qc;
      std::memset(&qc, 0, sizeof(qc));
      const int qcols = quad_pitch(c->cfg.kernel_size) - quad_u_pitch(c->cfg.kernel_size);
      qc.gridDim = dim3((unsigned)(1 + (qcols + kMmaCols - 1) / kMmaCols), (unsigned)((trn + 63) / 64));
      qc.blockDim = dim3(kMmaThreads);
      qc.dynamicSmemBytes = mma_smem_bytes(4, kTextStages);
      qc.stream = st;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      qa[0].val.programmaticStreamSerializationAllowed = 1;
      qc.attrs = qa;
      qc.numAttrs = pdl_ok() ? 1 : 0;
      CUDA_TRY(cudaLaunchKernelEx(&qc, quad_kernel, c->md, c->tb, tr0, trn));
      ++c->launches;
      prof_mark(c, "quad_kernel", st);
    }
  }
  // ---- K2 conv_image contraction with fused Find / stored FindSameProperty maps
  if (!S.work.empty()) {
    ProjParams p;
    p.work = reinterpret_cast<const ProjWork*>(d + o.work);
    p.num_work = (int)S.work.size();
    p.total_rows = c->md.N * c->HW;
    p.num_seg = c->md.num_seg;
    p.seg_images = c->md.N;
    p.n_tiles = c->Mp / 256;
    p.k_blocks = (c->Dk + kBK - 1) / kBK;
    p.HW = c->HW; p.M = c->cfg.map_dim; p.Mp = c->Mp; p.Dk = c->Dk;
    p.feat_pitch = c->md.feat_pitch;
    for (int s = 0; s < kMaxSeg; ++s) p.feat_seg[s] = c->md.feat_seg[s];
    for (int s = 0; s < NUM_PROJ_SETS; ++s) {
      p.bias[s] = c->proj_bias[s];
      p.w_orig[s] = c->md.proj_w[s];
    }
    p.img_ptr = reinterpret_cast<const int32_t*>(d + o.img_ptr);
    p.node_text = reinterpret_cast<const int32_t*>(d + o.node_text);
    p.node_out = reinterpret_cast<const int32_t*>(d + o.node_out);
    p.tauw = c->tb.tauw; p.tau2 = c->tb.tau2;
    p.elt_b = c->md.elt_b[ES_FIND];
    p.arena = arena;
    p.mslot = reinterpret_cast<const int32_t*>(d + o.mslot);
    p.num_images = (int)(S.mslot.size() / NUM_PROJ_SETS);
    p.mbuf = c->mbuf;
    // persistent grid of CTA pairs (clusters of 2): pair i walks work items i, i + pairs, ...
    const int max_ctas = c->proj_max_ctas > 0 ? std::min(c->proj_max_ctas, c->num_sms) : c->num_sms;
    const int pairs = std::max(1, std::min(p.num_work, max_ctas / 2));
    if (c->cfg.flags & N2NMN_FLAG_PROJ_FP32_SIMT) {
      const size_t smem = (size_t)(kSimtRows * kSimtKChunk + kSimtRows * c->Mp) * sizeof(float);
      proj_simt_kernel<<<2 * p.num_work, 256, smem, st>>>(p);
      prof_mark(c, "proj_simt_kernel", st);
    } else {
      cudaLaunchConfig_t lc;
      std::memset(&lc, 0, sizeof(lc));
      lc.gridDim = dim3((unsigned)(2 * pairs));
      lc.blockDim = dim3(kProjThreads);
      lc.dynamicSmemBytes = proj_smem_bytes(S.train);
      lc.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      lc.attrs = attr;
      lc.numAttrs = pdl_ok() ? 1 : 0;
      if (S.train) CUDA_TRY(cudaLaunchKernelEx(&lc, proj_umma_kernel<true>, c->tmaps, p));
      else CUDA_TRY(cudaLaunchKernelEx(&lc, proj_umma_kernel<false>, c->tmaps, p));
      prof_mark(c, "proj_umma_kernel", st);
    }
    ++c->launches;
  }
  // ---- K3 node evaluation
  NodeCtx nc;
  nc.md = c->md; nc.tb = c->tb; nc.arena = arena; nc.scores = scores_seg[0]; nc.mbuf = c->mbuf;
  nc.pooled = c->pooled; nc.pool_pitch = c->Kp; nc.pool_att = c->pool_att;
  nc.phi_out = S.train ? c->phi_buf : nullptr;
  nc.ehat = c->ehat; nc.ehat_lo = c->ehat_lo; nc.ehat_dst = c->ehat_dst;
  const int NQ = (int)S.q_ptr.size() - 1;
  // several segments: question q writes row q % N of segment q / N; one segment: row q (the
  // per-module entry point numbers its call rows beyond the bound batch size)
  const int nseg = std::max(1, S.num_seg);
  nc.score_rows = nseg > 1 ? S.N : (1 << 30);
  for (int s = 0; s < kMaxSeg; ++s) nc.scores_seg[s] = scores_seg[s < nseg ? s : 0];
  const bool ks3 = (c->cfg.kernel_size != 5);
  if (use_wave) {
    for (int s = 0; s < nseg; ++s)
      CUDA_TRY(cudaMemsetAsync(scores_seg[s], 0,
                               (size_t)(nseg > 1 ? S.N : NQ) * c->cfg.num_choices * sizeof(float),
                               st));
    const int32_t* d_wave = reinterpret_cast<const int32_t*>(d + o.wave_nodes);
    for (int dep = 1; dep <= S.max_depth; ++dep) {
      const int first = S.wave_ptr[dep], cnt = S.wave_ptr[dep + 1] - first;
      if (cnt == 0) continue;
      if (ks3) wave_kernel<3><<<cnt, kNodeThreads, c->node_smem_bytes, st>>>(nc, d_nodes, d_wave, first);
      else wave_kernel<5><<<cnt, kNodeThreads, c->node_smem_bytes, st>>>(nc, d_nodes, d_wave, first);
      ++c->launches;
      prof_mark(c, "wave_kernel", st);
    }
  } else if (NQ > 0) {
    // cluster size: spread one question over several SMs while the batch is small
    int cs = c->tree_cluster;
    if (cs <= 0) cs = (NQ * 4 <= 2 * c->num_sms) ? 4 : (NQ * 2 <= 2 * c->num_sms) ? 2 : 1;
    cudaLaunchConfig_t lc;
    std::memset(&lc, 0, sizeof(lc));
    lc.gridDim = dim3((unsigned)(NQ * cs));
    lc.blockDim = dim3(kNodeThreads);
    const int slots = std::max(1, S.max_stack);
    lc.dynamicSmemBytes = sizeof(float) * (size_t)tree_smem_layout(
        c->cfg.H, c->cfg.W, c->Mp, c->cfg.kernel_size, c->cfg.map_dim, c->cfg.num_choices,
        slots, S.pooled_direct).total;
    lc.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (cs > 1) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = (unsigned)cs;
      attr[na].val.clusterDim.y = 1;
      attr[na].val.clusterDim.z = 1;
      ++na;
    }
    if (pdl_ok()) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    lc.attrs = attr;
    lc.numAttrs = na;
    const int wa = (write_arena ? kTreeWriteArena : 0) |
                   (((c->cfg.flags & N2NMN_FLAG_PROJ_FP32_SIMT) || c->fp32_stencil) ? kTreeFp32Stencil : 0);
    if (S.pooled_direct) {
      if (ks3) CUDA_TRY(cudaLaunchKernelEx(&lc, tree_kernel<3, true>, nc, d_nodes, d_qptr, cs, slots, wa));
      else CUDA_TRY(cudaLaunchKernelEx(&lc, tree_kernel<5, true>, nc, d_nodes, d_qptr, cs, slots, wa));
    } else {
      if (ks3) CUDA_TRY(cudaLaunchKernelEx(&lc, tree_kernel<3, false>, nc, d_nodes, d_qptr, cs, slots, wa));
      else CUDA_TRY(cudaLaunchKernelEx(&lc, tree_kernel<5, false>, nc, d_nodes, d_qptr, cs, slots, wa));
    }
    ++c->launches;
    prof_mark(c, "tree_kernel", st);
    // ---- K4 batched answer heads of the attention-pooled roots
    if (S.pooled_direct && !S.head_work.empty()) {
      if ((int)S.num_pool_rows > 2 * c->QB)
        return fail(N2NMN_ERR_CAPACITY, "too many pooled root nodes for this context");
      cudaLaunchConfig_t hc;
      std::memset(&hc, 0, sizeof(hc));
      hc.gridDim = dim3((unsigned)S.head_work.size());
      hc.blockDim = dim3(kHeadThreads);
      hc.dynamicSmemBytes = (size_t)c->head_smem_bytes;
      hc.stream = st;
      cudaLaunchAttribute hattr[1];
      hattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      hattr[0].val.programmaticStreamSerializationAllowed = 1;
      hc.attrs = hattr;
      hc.numAttrs = pdl_ok() ? 1 : 0;
      if (S.num_feat_rows > 0) {   // pooled features: one CTA per (root row, 128-channel chunk)
        const int HWp = (c->HW + 3) & ~3;
        cudaLaunchConfig_t pc = hc;
        const int quads = c->md.feat_pitch / 4;
        pc.gridDim = dim3((unsigned)S.num_feat_rows, (unsigned)((quads + kPoolQuads - 1) / kPoolQuads));
        pc.blockDim = dim3(kPoolQuads * kPoolSlices);
        pc.dynamicSmemBytes = (size_t)(HWp + 4 * kPoolQuads * kPoolSlices) * sizeof(float);
        CUDA_TRY(cudaLaunchKernelEx(&pc, pool_kernel, nc,
                                    reinterpret_cast<const int32_t*>(d + o.pool_img), HWp));
        ++c->launches;
        prof_mark(c, "pool_kernel", st);
      }
      const HeadWork* d_hw = reinterpret_cast<const HeadWork*>(d + o.head_work);
      const int32_t* d_hl = reinterpret_cast<const int32_t*>(d + o.head_list);
      if (c->head_nn == 16) CUDA_TRY(cudaLaunchKernelEx(&hc, head_kernel<16>, nc, d_nodes, d_hw, d_hl));
      else if (c->head_nn == 8) CUDA_TRY(cudaLaunchKernelEx(&hc, head_kernel<8>, nc, d_nodes, d_hw, d_hl));
      else CUDA_TRY(cudaLaunchKernelEx(&hc, head_kernel<4>, nc, d_nodes, d_hw, d_hl));
      ++c->launches;
      prof_mark(c, "head_kernel", st);
      if (c->ehat) {   // fc_eltwise of the Describe-type roots as one GEMM per weight set
        for (int op : {OP_DESCRIBE, OP_SAME_PROPERTY}) {
          int r0 = 1 << 30, r1 = 0;
          for (const HeadWork& w : S.head_work)
            if (w.op == op) { r0 = std::min(r0, (int)w.first); r1 = std::max(r1, (int)(w.first + w.count)); }
          if (r1 <= r0) continue;
          const int os = op == OP_DESCRIBE ? OS_DESCRIBE : OS_SAMEPROP;
          if (!c->out_wp[os]) return fail(N2NMN_ERR_STATE, "answer-head weights not packed");
          if (c->ehat_lo && c->out_wt_hi[os] && !std::getenv("N2NMN_TAIL_MMA_SYNC")) {
            cudaLaunchConfig_t tc = hc;
            tc.gridDim = dim3((unsigned)(c->Cpad / kHtN), (unsigned)((r1 - r0 + kHtM - 1) / kHtM));
            tc.blockDim = dim3(kHtThreads);
            tc.dynamicSmemBytes = kHtSmemBytes;
            CUDA_TRY(cudaLaunchKernelEx(&tc, head_tail_umma_kernel, c->ht_maps[os], c->md.out_b[os],
                                        (float* const*)c->ehat_dst, r0, r1 - r0,
                                        (int)c->cfg.num_choices, (int)c->cfg.map_dim));
            ++c->launches;
            prof_mark(c, "head_tail_gemm_kernel", st);
            continue;
          }
          GemmOperands gp;
          gp.a0 = c->ehat + (size_t)r0 * c->Mp; gp.k0 = c->cfg.map_dim; gp.lda0 = c->Mp;
          gp.a1 = nullptr; gp.k1 = 0; gp.lda1 = 0;
          gp.R = r1 - r0; gp.B = c->out_wp[os]; gp.ldb = c->Cp; gp.C = c->cfg.num_choices;
          cudaLaunchConfig_t gc = hc;
          const bool narrow = gp.R <= 32;
          gc.gridDim = dim3((unsigned)((gp.C + kMmaCols - 1) / kMmaCols),
                            (unsigned)(narrow ? (gp.R + 31) / 32 : (gp.R + 63) / 64));
          gc.blockDim = dim3(kMmaThreads);
          gc.dynamicSmemBytes = mma_smem_bytes(narrow ? 2 : 4);
          if (narrow) CUDA_TRY(cudaLaunchKernelEx(&gc, head_tail_gemm_kernel<2>, gp, c->md.out_b[os],
                                                  (float* const*)c->ehat_dst, r0));
          else CUDA_TRY(cudaLaunchKernelEx(&gc, head_tail_gemm_kernel<4>, gp, c->md.out_b[os],
                                           (float* const*)c->ehat_dst, r0));
          ++c->launches;
          prof_mark(c, "head_tail_gemm_kernel", st);
        }
      }
    }
  }
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(slot->last_use, st));
  return 0;
}

int check_ready(n2nmn_ctx* c) {
  if (!c->bound) return fail(N2NMN_ERR_STATE, "n2nmn_bind_inputs has not been called");
  for (const Variable& v : c->vars)
    if (!v.loaded) return fail(N2NMN_ERR_STATE, "variable not set: " + v.name);
  return 0;
}

}  // namespace

// ================================================================================== C ABI
namespace n2nmn {
int fail_with(int code, const std::string& msg) { return fail(code, msg); }   // other TUs
}

extern "C" {

const char* n2nmn_last_error(void) { return g_err.c_str(); }

int n2nmn_create(const n2nmn_config* cfg, n2nmn_ctx** out) {
  if (!cfg || !out) return fail(N2NMN_ERR_ARG, "null argument");
  if (cfg->abi_version != N2NMN_ABI_VERSION) return fail(N2NMN_ERR_ARG, "ABI version mismatch");
  if (cfg->family < 0 || cfg->family > 2 || cfg->H <= 0 || cfg->W <= 0 || cfg->D <= 0 ||
      cfg->map_dim <= 0 || cfg->map_dim > 1024 || cfg->num_choices <= 0 || cfg->max_batch <= 0 ||
      cfg->max_T <= 0 || cfg->text_dim <= 0 || cfg->max_group < 0 || cfg->max_group > kMaxSeg)
    return fail(N2NMN_ERR_ARG, "bad configuration");
  if (cfg->family != N2NMN_VQA && cfg->kernel_size != 3 && cfg->kernel_size != 5)
    return fail(N2NMN_ERR_ARG, "kernel_size must be 3 or 5");
  CUDA_TRY(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10)
    return fail(N2NMN_ERR_DEVICE, std::string("n2nmn_b200 needs an sm_100 GPU, found sm_") +
                                      std::to_string(prop.major) + std::to_string(prop.minor));
  n2nmn_ctx* c = new n2nmn_ctx();
  c->cfg = *cfg;
  if (cfg->family == N2NMN_VQA) c->cfg.kernel_size = 1;
  c->device = cfg->device;
  c->num_sms = prop.multiProcessorCount;
  c->HW = cfg->H * cfg->W;
  c->Dk = cfg->D + (cfg->family == N2NMN_VQA ? 2 : 0);
  c->Kp = round_up(c->Dk, kBK);
  c->Mp = round_up(cfg->map_dim, 256);
  c->G = std::max(1, std::min(cfg->max_group, kMaxSeg));
  c->QB = c->G * cfg->max_batch;
  std::memset(&c->md, 0, sizeof(c->md));
  DevModel& md = c->md;
  md.H = cfg->H; md.W = cfg->W; md.HW = c->HW; md.Dk = c->Dk; md.Dt = cfg->text_dim;
  md.M = cfg->map_dim; md.Mp = c->Mp; md.C = cfg->num_choices; md.ksize = c->cfg.kernel_size;
  md.family = cfg->family;
  c->shp = SchedShape{cfg->family, cfg->H, cfg->W, c->Dk, cfg->text_dim, cfg->map_dim, c->Mp,
                      cfg->num_choices, c->cfg.kernel_size, cfg->max_T};
  build_variables(c);
  {   // flat TF-layout buffer (weights / gradients / Adam moments) and where each gradient goes
    int64_t off = 0;
    std::memset(&c->go, 0, sizeof(c->go));
    for (const Variable& v : c->vars) {
      c->flat_offset.push_back(off);
      const int o = (int)off;
      if (v.kind == VK_PROJ_W) c->go.proj_w[v.set] = o;
      else if (v.kind == VK_PROJ_B) c->go.proj_b[v.set] = o;
      else {
        for (int i = 0; i < NUM_TEXT_SETS; ++i) {
          if (v.slot == &md.txt_w[i]) c->go.txt_w[i] = o;
          if (v.slot == &md.txt_b[i]) c->go.txt_b[i] = o;
        }
        for (int i = 0; i < NUM_ELT_SETS; ++i) {
          if (v.slot == &md.elt_w[i]) c->go.elt_w[i] = o;
          if (v.slot == &md.elt_b[i]) c->go.elt_b[i] = o;
        }
        for (int i = 0; i < NUM_OUT_SETS; ++i) {
          if (v.slot == &md.out_w[i]) c->go.out_w[i] = o;
          if (v.slot == &md.out_b[i]) c->go.out_b[i] = o;
        }
        for (int i = 0; i < NUM_SCORE_SETS; ++i) {
          if (v.slot == &md.sc_w[i]) c->go.sc_w[i] = o;
          if (v.slot == &md.sc_b[i]) c->go.sc_b[i] = o;
        }
        if (v.slot == &md.conv_k) c->go.conv_k = o;
        if (v.slot == &md.conv_b) c->go.conv_b = o;
      }
      off += (int64_t)((v.count + 3) & ~(size_t)3);
    }
    c->flat_size = off;
  }

  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess)
    return fail(N2NMN_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  c->encode = reinterpret_cast<EncodeTiledFn>(fn);

  // weights: plain copies + repacked projection operands
  CUDA_TRY(cudaMalloc(&c->wbuf, c->wbuf_floats * sizeof(float)));
  CUDA_TRY(cudaMemset(c->wbuf, 0, c->wbuf_floats * sizeof(float)));
  for (Variable& v : c->vars)
    if (v.slot) *v.slot = c->wbuf + v.offset;
  for (const Variable& v : c->vars)
    if (v.kind == VK_PROJ_W) c->proj_used[v.set] = true;
  for (int s = 0; s < NUM_PROJ_SETS; ++s) {
    if (!c->proj_used[s]) continue;
    if (s != PS_FIND) ++c->num_store_sets;
    CUDA_TRY(cudaMalloc(&c->proj_wt[s], (size_t)c->Mp * c->Kp * sizeof(float)));
    CUDA_TRY(cudaMemset(c->proj_wt[s], 0, (size_t)c->Mp * c->Kp * sizeof(float)));
    CUDA_TRY(cudaMalloc(&c->proj_bias[s], (size_t)c->Mp * sizeof(float)));
    CUDA_TRY(cudaMemset(c->proj_bias[s], 0, (size_t)c->Mp * sizeof(float)));
    md.proj_b[s] = c->proj_bias[s];
    if (int rc = encode_2d(c, &c->tmaps.b[s], c->proj_wt[s], c->Kp, c->Mp, c->Kp, kBK, kBNHalf))
      return rc;
  }
  for (int s = 0; s < NUM_PROJ_SETS; ++s)      // sets the family lacks alias set 0 (never used)
    if (!c->proj_used[s]) c->tmaps.b[s] = c->tmaps.b[PS_FIND];
  // workspaces (sized for a full group of segments)
  const int NB = c->QB, TT = cfg->max_T;
  c->text_rows_cap = NB * TT;
  const size_t tb_floats = (size_t)c->text_rows_cap * c->Mp;
  CUDA_TRY(cudaMalloc(&c->tb.tau, 3 * tb_floats * sizeof(float)));
  c->tb.tauw = c->tb.tau + tb_floats;
  c->tb.tau2 = c->tb.tauw + tb_floats;
  c->arena_slots = std::max(NB * TT, 3 * NB);
  CUDA_TRY(cudaMalloc(&c->arena, (size_t)c->arena_slots * c->HW * sizeof(float)));
  c->mbuf_slots = (c->num_store_sets + 1) * NB;   // +1: conv_image maps of Find kept for backward
  CUDA_TRY(cudaMalloc(&c->mbuf, (size_t)c->mbuf_slots * c->HW * c->Mp * sizeof(float)));
  CUDA_TRY(cudaMalloc(&c->scores_tmp,
                      (size_t)cfg->max_batch * TT * cfg->num_choices * sizeof(float)));
  CUDA_TRY(cudaMalloc(&c->pooled, (size_t)2 * NB * c->Kp * sizeof(float)));
  CUDA_TRY(cudaMalloc(&c->pool_att, (size_t)2 * NB * ((c->HW + 3) & ~3) * sizeof(float)));
  if (c->cfg.family != N2NMN_VQA) {
    const int ks = c->cfg.kernel_size;
    CUDA_TRY(cudaMalloc(&c->conv_quad, (size_t)quad_pitch(ks) * c->Mp * sizeof(float)));
    CUDA_TRY(cudaMemset(c->conv_quad, 0, (size_t)quad_pitch(ks) * c->Mp * sizeof(float)));
    CUDA_TRY(cudaMalloc(&c->tb.tq, (size_t)c->text_rows_cap * quad_pitch(ks) * sizeof(float)));
    md.conv_quad = c->conv_quad;
    if (quad_pitch(ks) > 3 * c->Mp)
      return fail(N2NMN_ERR_ARG, "kernel_size too large for this map_dim");
  }
  if (2 * (c->HW + 2) * kHeadNodesMax > head_smem_floats(c->head_nn, c->Kp, c->Mp) - 8 * kHeadNodesMax * 32)
    return fail(N2NMN_ERR_ARG, "grid too large for the answer-head kernel");
  if (cfg->num_choices > 32) {
    c->Cp = round_up(cfg->num_choices, 4);
    CUDA_TRY(cudaMalloc(&c->ehat, (size_t)NB * c->Mp * sizeof(float)));
    CUDA_TRY(cudaMalloc(&c->ehat_dst, (size_t)NB * sizeof(float*)));
    c->Cpad = round_up(cfg->num_choices, kHtN);
    CUDA_TRY(cudaMalloc(&c->ehat_lo, (size_t)NB * c->Mp * sizeof(float)));
    for (const Variable& v : c->vars)
      for (int os = 0; os < NUM_OUT_SETS; ++os)
        if (v.slot == &md.out_w[os]) {
          CUDA_TRY(cudaMalloc(&c->out_wp[os], (size_t)cfg->map_dim * c->Cp * sizeof(float)));
          CUDA_TRY(cudaMemset(c->out_wp[os], 0, (size_t)cfg->map_dim * c->Cp * sizeof(float)));
          const size_t wt = (size_t)c->Cpad * c->Mp * sizeof(float);
          CUDA_TRY(cudaMalloc(&c->out_wt_hi[os], wt));
          CUDA_TRY(cudaMalloc(&c->out_wt_lo[os], wt));
          CUDA_TRY(cudaMemset(c->out_wt_hi[os], 0, wt));
          CUDA_TRY(cudaMemset(c->out_wt_lo[os], 0, wt));
          HeadTailMaps& hm = c->ht_maps[os];
          if (int rc = encode_2d(c, &hm.a_hi, c->ehat, c->Mp, NB, c->Mp, kHtK, kHtM)) return rc;
          if (int rc = encode_2d(c, &hm.a_lo, c->ehat_lo, c->Mp, NB, c->Mp, kHtK, kHtM)) return rc;
          if (int rc = encode_2d(c, &hm.b_hi, c->out_wt_hi[os], c->Mp, c->Cpad, c->Mp, kHtK, kHtN)) return rc;
          if (int rc = encode_2d(c, &hm.b_lo, c->out_wt_lo[os], c->Mp, c->Cpad, c->Mp, kHtK, kHtN)) return rc;
        }
    CUDA_TRY(cudaFuncSetAttribute(head_tail_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kHtSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(head_tail_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)mma_smem_bytes(2)));
    CUDA_TRY(cudaFuncSetAttribute(head_tail_gemm_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)mma_smem_bytes(4)));
  }
  c->head_nn = head_nodes_per_cta(c->Dk, c->Mp);
  c->head_smem_bytes = head_smem_layout(c->head_nn, c->Kp, c->Mp).total * (int)sizeof(float);
  if (cfg->family == N2NMN_VQA || (cfg->D % 4) != 0) {
    CUDA_TRY(cudaMalloc(&c->feat_aug, (size_t)NB * c->HW * c->Kp * sizeof(float)));
  }
  // schedule tables: generous upper bound on every table for (max_batch, max_T)
  {
    const size_t nodes = (size_t)NB * TT;
    const size_t tiles = (size_t)c->G * (((size_t)cfg->max_batch * c->HW + 127) / 128 + 1);
    c->table_cap = nodes * (sizeof(NodeRec) + 4 * 6) + (nodes / 8 + 8) * sizeof(TextGroup) +
                   tiles * (TT / kMaxProjNodesPerPass + 1 + NUM_PROJ_SETS) * sizeof(ProjWork) +
                   (size_t)NB * (12 + 4 * NUM_PROJ_SETS) + nodes * (16 + 2 * sizeof(BwdEntryHost)) +
                   (size_t)NB * (4 + 8 + sizeof(HeadWork)) + 4096;
    for (int i = 0; i < kTableSlots; ++i) {
      CUDA_TRY(cudaMallocHost(&c->slots[i].host, c->table_cap));
      CUDA_TRY(cudaMalloc(&c->slots[i].dev, c->table_cap));
      CUDA_TRY(cudaEventCreateWithFlags(&c->slots[i].last_use, cudaEventDisableTiming));
    }
  }
  // kernel attributes
  const NodeSmem L = node_smem_layout(cfg->H, cfg->W, c->Mp, c->cfg.kernel_size, cfg->map_dim,
                                      cfg->num_choices);
  c->node_smem_bytes = L.total * (int)sizeof(float);
  c->stack_cap = cfg->max_T / 2 + 2;
  for (;;) {   // largest attention stack that still fits in shared memory
    c->tree_smem_bytes = (int)sizeof(float) * tree_smem_layout(
        cfg->H, cfg->W, c->Mp, c->cfg.kernel_size, cfg->map_dim, cfg->num_choices,
        c->stack_cap).total;
    if (c->tree_smem_bytes <= 200 * 1024 || c->stack_cap <= 2) break;
    --c->stack_cap;
  }
  CUDA_TRY(cudaFuncSetAttribute(tree_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                c->tree_smem_bytes));
  CUDA_TRY(cudaFuncSetAttribute(tree_kernel<5, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                c->tree_smem_bytes));
  CUDA_TRY(cudaFuncSetAttribute(tree_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                c->tree_smem_bytes));
  CUDA_TRY(cudaFuncSetAttribute(tree_kernel<5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                c->tree_smem_bytes));
  {
    if (cfg->text_dim % 4 != 0)
      return fail(N2NMN_ERR_ARG, "text_dim must be a multiple of 4 (16-byte word-vector rows)");
    CUDA_TRY(cudaFuncSetAttribute(text_proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)mma_smem_bytes(4, kTextStages)));
    CUDA_TRY(cudaFuncSetAttribute(quad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)mma_smem_bytes(4, kTextStages)));
    CUDA_TRY(cudaFuncSetAttribute(text_proj_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CUDA_TRY(cudaFuncSetAttribute(quad_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  }
  CUDA_TRY(cudaFuncSetAttribute(head_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                head_smem_layout(16, c->Kp, c->Mp).total * (int)sizeof(float) <=
                                        200 * 1024
                                    ? head_smem_layout(16, c->Kp, c->Mp).total * (int)sizeof(float)
                                    : 48 * 1024));
  CUDA_TRY(cudaFuncSetAttribute(head_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                head_smem_layout(8, c->Kp, c->Mp).total * (int)sizeof(float) <=
                                        200 * 1024
                                    ? head_smem_layout(8, c->Kp, c->Mp).total * (int)sizeof(float)
                                    : 48 * 1024));
  CUDA_TRY(cudaFuncSetAttribute(head_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                head_smem_layout(4, c->Kp, c->Mp).total * (int)sizeof(float) <=
                                        200 * 1024
                                    ? head_smem_layout(4, c->Kp, c->Mp).total * (int)sizeof(float)
                                    : 48 * 1024));
  if (c->head_smem_bytes > 200 * 1024)
    return fail(N2NMN_ERR_ARG, "feature depth too large for the answer-head kernel");
  CUDA_TRY(cudaFuncSetAttribute(wave_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                c->node_smem_bytes));
  CUDA_TRY(cudaFuncSetAttribute(wave_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                c->node_smem_bytes));
  CUDA_TRY(cudaFuncSetAttribute(proj_umma_kernel<true>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, proj_smem_bytes(true)));
  CUDA_TRY(cudaFuncSetAttribute(proj_umma_kernel<false>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                proj_smem_bytes(false)));
  CUDA_TRY(cudaFuncSetAttribute(
      proj_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
      (int)((kSimtRows * kSimtKChunk + kSimtRows * c->Mp) * sizeof(float))));
  c->module_sched.uid = g_uid++;
  c->module_sched.shp = c->shp;
  c->step_sched.shp = c->shp;
  if (const char* e = std::getenv("N2NMN_NO_PDL")) c->use_pdl = (std::atoi(e) == 0);
  if (const char* e = std::getenv("N2NMN_FP32_STENCIL")) c->fp32_stencil = (std::atoi(e) != 0);
  if (const char* e = std::getenv("N2NMN_TEXT_CTAS")) c->text_ctas_per_group = std::max(0, std::atoi(e));
  if (const char* e = std::getenv("N2NMN_PROJ_CTAS")) c->proj_max_ctas = std::max(0, std::atoi(e));
  if (const char* e = std::getenv("N2NMN_TREE_CLUSTER")) {
    const int v = std::atoi(e);
    if (v == 1 || v == 2 || v == 4 || v == 8) c->tree_cluster = v;
  }
  *out = c;
  return 0;
}

int n2nmn_destroy(n2nmn_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  cudaFree(c->wbuf);
  for (int s = 0; s < NUM_PROJ_SETS; ++s) { cudaFree(c->proj_wt[s]); cudaFree(c->proj_bias[s]); }
  cudaFree(c->feat_aug); cudaFree(c->tb.tau); cudaFree(c->arena); cudaFree(c->mbuf);
  cudaFree(c->pooled); cudaFree(c->pool_att); cudaFree(c->conv_quad); cudaFree(c->tb.tq);
  cudaFree(c->dscores); cudaFree(c->per_sample); cudaFree(c->dtau); cudaFree(c->dmap); cudaFree(c->dstencil); cudaFree(c->gmap); cudaFree(c->phi_buf);
  cudaFree(c->ehat); cudaFree(c->ehat_dst); cudaFree(c->ehat_lo);
  for (int os = 0; os < NUM_OUT_SETS; ++os) {
    cudaFree(c->out_wp[os]); cudaFree(c->out_wt_hi[os]); cudaFree(c->out_wt_lo[os]);
  }
  cudaFree(c->d_segs); cudaFree(c->d_sumsq);
  cudaFree(c->scores_tmp); cudaFree(c->e2e_feat); cudaFree(c->e2e_wv); cudaFree(c->e2e_scores);
  cudaFree(c->e2e_feat_f16);
  for (int i = 0; i < kTableSlots; ++i) {
    cudaFreeHost(c->slots[i].host); cudaFree(c->slots[i].dev);
    if (c->slots[i].last_use) cudaEventDestroy(c->slots[i].last_use);
  }
  for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
  delete c;
  return 0;
}

int n2nmn_num_variables(const n2nmn_ctx* c) { return c ? (int)c->vars.size() : 0; }

int n2nmn_variable_info(const n2nmn_ctx* c, int index, const char** name, int64_t shape[4],
                        int* ndim) {
  if (!c || index < 0 || index >= (int)c->vars.size()) return fail(N2NMN_ERR_ARG, "bad index");
  const Variable& v = c->vars[index];
  if (name) *name = v.name.c_str();
  if (ndim) *ndim = (int)v.shape.size();
  if (shape) for (size_t i = 0; i < v.shape.size() && i < 4; ++i) shape[i] = v.shape[i];
  return 0;
}

int n2nmn_set_weight(n2nmn_ctx* c, const char* name, const float* src, const int64_t* shape,
                     int ndim, void* stream) {
  if (!c || !name || !src) return fail(N2NMN_ERR_ARG, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (Variable& v : c->vars) {
    if (v.name != name) continue;
    if (ndim != (int)v.shape.size()) return fail(N2NMN_ERR_ARG, std::string("rank mismatch for ") + name);
    for (int i = 0; i < ndim; ++i)
      if (shape[i] != v.shape[i]) return fail(N2NMN_ERR_ARG, std::string("shape mismatch for ") + name);
    if (v.kind == VK_PITCHED) {
      const int rows = (int)(v.count / (size_t)v.shape.back());
      pitch_rows_kernel<<<(unsigned)rows, 256, 0, st>>>(src, rows, (int)v.shape.back(),
                                                      c->wbuf + v.offset, c->Mp);
    } else {
      CUDA_TRY(cudaMemcpyAsync(c->wbuf + v.offset, src, v.count * sizeof(float),
                               cudaMemcpyDeviceToDevice, st));
    }
    if (v.kind == VK_PROJ_W) {
      dim3 grid((c->Kp + 31) / 32, (c->Mp + 31) / 32), block(32, 8);
      transpose_pad_kernel<<<grid, block, 0, st>>>(c->wbuf + v.offset, c->Dk, c->cfg.map_dim,
                                                   c->proj_wt[v.set], c->Kp, c->Mp);
    } else if (v.kind == VK_PROJ_B) {
      pad_copy_kernel<<<(c->Mp + 255) / 256, 256, 0, st>>>(c->wbuf + v.offset, c->cfg.map_dim,
                                                           c->proj_bias[v.set], c->Mp);
    }
    for (int os = 0; os < NUM_OUT_SETS; ++os)
      if (v.slot == &c->md.out_w[os] && c->out_wp[os]) {
        pitch_rows_kernel<<<(unsigned)c->cfg.map_dim, 256, 0, st>>>(src, c->cfg.map_dim,
                                                                   c->cfg.num_choices, c->out_wp[os], c->Cp);
        if (c->out_wt_hi[os])
          out_wt_split_kernel<<<dim3((c->Cpad + 31) / 32, (c->Mp + 31) / 32), dim3(32, 8), 0, st>>>(
              src, c->cfg.map_dim, c->cfg.num_choices, c->out_wt_hi[os], c->out_wt_lo[os], c->Mp,
              c->Cpad);
      }
    if (c->conv_quad && (v.slot == &c->md.conv_k || v.slot == &c->md.conv_b ||
                         v.slot == &c->md.elt_w[ES_TRANSFORM])) {
      // the Transform quadratic-form matrix depends on these three variables
      conv_quad_kernel<<<quad_pitch(c->cfg.kernel_size), 256, 0, st>>>(
          c->md.conv_k, c->md.conv_b, c->md.elt_w[ES_TRANSFORM], c->cfg.kernel_size,
          c->cfg.map_dim, c->Mp, c->conv_quad);
    }
    CUDA_TRY(cudaGetLastError());
    v.loaded = true;
    return 0;
  }
  return fail(N2NMN_ERR_ARG, std::string("unknown variable: ") + name);
}

namespace {
// Points the context at `nseg` batches of identical shape (segments): feature grids [N,H,W,D],
// word vectors [T,N,Dt]. One TMA tensor map per segment; VQA / odd channel counts get their
// augmented / re-pitched copy per segment.
int bind_segments(n2nmn_ctx* c, int nseg, const float* const* feat, const float* const* wv, int N,
                  int T, cudaStream_t st) {
  if (nseg <= 0 || nseg > c->G) return fail(N2NMN_ERR_CAPACITY, "too many batches for one group");
  if (N <= 0 || T <= 0) return fail(N2NMN_ERR_ARG, "N and T must be positive");
  if (N > c->cfg.max_batch || T > c->cfg.max_T)
    return fail(N2NMN_ERR_CAPACITY, "N or T exceeds the context capacity");
  const int rows = N * c->HW;
  const int pitch = c->feat_aug ? c->Kp : c->cfg.D;
  for (int sgi = 0; sgi < nseg; ++sgi) {
    if (!feat[sgi] || !wv[sgi]) return fail(N2NMN_ERR_ARG, "null argument");
    const float* eff = feat[sgi];
    if (c->feat_aug) {
      float* dst = c->feat_aug + (size_t)sgi * c->cfg.max_batch * c->HW * c->Kp;
      augment_features_kernel<<<rows, 128, 0, st>>>(feat[sgi], rows, c->cfg.D, c->cfg.H, c->cfg.W,
                                                   c->cfg.family == N2NMN_VQA ? 1 : 0, dst, c->Kp);
      CUDA_TRY(cudaGetLastError());
      ++c->launches;
      eff = dst;
    }
    if ((reinterpret_cast<uintptr_t>(eff) & 15) != 0)
      return fail(N2NMN_ERR_ARG, "image_feat_grid must be 16-byte aligned");
    c->md.feat_seg[sgi] = eff;
    c->md.wv_seg[sgi] = wv[sgi];
    if (int rc = encode_2d(c, &c->tmaps.a[sgi], eff, c->Dk, rows, pitch, kBK, kBM)) return rc;
  }
  for (int sgi = nseg; sgi < kMaxSeg; ++sgi) {
    c->md.feat_seg[sgi] = c->md.feat_seg[0];
    c->md.wv_seg[sgi] = c->md.wv_seg[0];
  }
  c->md.feat = c->md.feat_seg[0]; c->md.feat_pitch = pitch; c->md.word_vecs = c->md.wv_seg[0];
  c->md.N = N; c->md.T = T; c->md.num_seg = nseg;
  c->N = N; c->T = T;
  c->bound = true;
  return 0;
}
}  // namespace

int n2nmn_bind_inputs(n2nmn_ctx* c, const float* feat, const float* wv, int N, int T,
                      void* stream) {
  if (!c || !feat || !wv) return fail(N2NMN_ERR_ARG, "null argument");
  return bind_segments(c, 1, &feat, &wv, N, T, static_cast<cudaStream_t>(stream));
}

int n2nmn_compile_schedule(n2nmn_ctx* c, const int32_t* tokens, int T, int N,
                           const int32_t* vocab_ops, int num_vocab, uint8_t* validity_out,
                           n2nmn_sched** out) {
  if (!c || !tokens || !vocab_ops || !out) return fail(N2NMN_ERR_ARG, "null argument");
  if (N <= 0 || T <= 0) return fail(N2NMN_ERR_ARG, "N and T must be positive");
  if (N > c->cfg.max_batch || T > c->cfg.max_T)
    return fail(N2NMN_ERR_CAPACITY, "N or T exceeds the context capacity");
  n2nmn_sched* sc = new n2nmn_sched();
  sc->uid = g_uid++;
  sc->shp = c->shp;
  const char* err = nullptr;
  const bool direct = !(c->cfg.flags & N2NMN_FLAG_WAVE_EXECUTOR);
  const int rc = compile_schedule_group(c->shp, &tokens, 1, T, N, vocab_ops, num_vocab, &sc->hs,
                                        &err, false, direct);
  if (rc) { delete sc; return fail(rc, err ? err : "compile_schedule failed"); }
  if (validity_out) std::memcpy(validity_out, sc->hs.validity.data(), N);
  *out = sc;
  return 0;
}

int n2nmn_compile_schedule_host(const n2nmn_config* cfg, const int32_t* tokens, int T, int N,
                                const int32_t* vocab_ops, int num_vocab, uint8_t* validity_out,
                                n2nmn_sched** out) {
  if (!cfg || !tokens || !vocab_ops || !out) return fail(N2NMN_ERR_ARG, "null argument");
  if (N <= 0 || T <= 0) return fail(N2NMN_ERR_ARG, "N and T must be positive");
  const int Dk = cfg->D + (cfg->family == N2NMN_VQA ? 2 : 0);
  const SchedShape shp{cfg->family, cfg->H, cfg->W, Dk, cfg->text_dim, cfg->map_dim,
                       round_up(cfg->map_dim, 256), cfg->num_choices,
                       cfg->family == N2NMN_VQA ? 1 : cfg->kernel_size, cfg->max_T};
  n2nmn_sched* sc = new n2nmn_sched();
  sc->uid = g_uid++;
  sc->shp = shp;
  const char* err = nullptr;
  const int rc = compile_schedule(shp, tokens, T, N, vocab_ops, num_vocab, &sc->hs, &err);
  if (rc) { delete sc; return fail(rc, err ? err : "compile_schedule failed"); }
  if (validity_out) std::memcpy(validity_out, sc->hs.validity.data(), N);
  *out = sc;
  return 0;
}

// Diagnostic: nanoseconds per layout compile (host only), reusing one schedule object the way
// n2nmn_forward_tokens does.
double n2nmn_time_compile(const n2nmn_config* cfg, const int32_t* tokens, int T, int N,
                          const int32_t* vocab_ops, int num_vocab, int iters) {
  if (!cfg || !tokens || !vocab_ops || iters <= 0) return -1.0;
  const int Dk = cfg->D + (cfg->family == N2NMN_VQA ? 2 : 0);
  const SchedShape shp{cfg->family, cfg->H, cfg->W, Dk, cfg->text_dim, cfg->map_dim,
                       round_up(cfg->map_dim, 256), cfg->num_choices,
                       cfg->family == N2NMN_VQA ? 1 : cfg->kernel_size, cfg->max_T};
  HostSchedule hs;
  const char* err = nullptr;
  compile_schedule(shp, tokens, T, N, vocab_ops, num_vocab, &hs, &err);
  timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int i = 0; i < iters; ++i) compile_schedule(shp, tokens, T, N, vocab_ops, num_vocab, &hs, &err);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return ((t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec)) / iters;
}

int n2nmn_compile_nodes(n2nmn_ctx* c, const int32_t* op, const int32_t* t_idx,
                        const int32_t* b_idx, const int32_t* in0, const int32_t* in1, int n,
                        const int32_t* q_ptr, int nq, n2nmn_sched** out) {
  if (!c || !q_ptr || !out || (n > 0 && (!op || !t_idx || !b_idx || !in0 || !in1)))
    return fail(N2NMN_ERR_ARG, "null argument");
  if (nq <= 0 || nq > c->cfg.max_batch) return fail(N2NMN_ERR_CAPACITY, "bad question count");
  if (q_ptr[0] != 0 || q_ptr[nq] != n) return fail(N2NMN_ERR_ARG, "q_ptr does not span the nodes");
  static const int arity[NUM_OPS] = {0, 0, 1, 1, 1, 2, 2, 1, 1, 2, 2, 2, 2, 1};
  static const bool is_ans[NUM_OPS] = {0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1};
  n2nmn_sched* sc = new n2nmn_sched();
  sc->uid = g_uid++;
  sc->shp = c->shp;
  HostSchedule& S = sc->hs;
  S.N = c->cfg.max_batch; S.T = c->cfg.max_T;   // batch_idx may name any bound image
  S.pooled_direct = !(c->cfg.flags & N2NMN_FLAG_WAVE_EXECUTOR);
  S.nodes.resize(n); S.depth.assign(n, 1);
  S.q_ptr.assign(q_ptr, q_ptr + nq + 1);
  S.validity.assign(nq, 0);
  const int scene_bits = [] { float v = 3.0f; int b; std::memcpy(&b, &v, 4); return b; }();
  for (int q = 0; q < nq; ++q) {
    if (q_ptr[q + 1] < q_ptr[q]) { delete sc; return fail(N2NMN_ERR_ARG, "q_ptr not monotone"); }
    if (q_ptr[q + 1] > q_ptr[q]) { S.validity[q] = 1; ++S.num_valid; }
    for (int i = q_ptr[q]; i < q_ptr[q + 1]; ++i) {
      NodeRec& r = S.nodes[i];
      const int o = op[i];
      bool ok = o >= 0 && o < NUM_OPS && t_idx[i] >= 0 && t_idx[i] < c->cfg.max_T &&
                b_idx[i] >= 0 && b_idx[i] < c->cfg.max_batch;
      const int kids[2] = {in0[i], in1[i]};
      for (int k = 0; ok && k < 2; ++k) {
        if (k < arity[o]) {
          ok = kids[k] >= q_ptr[q] && kids[k] < i && !is_ans[op[kids[k]]];
          if (ok) S.depth[i] = std::max(S.depth[i], S.depth[kids[k]] + 1);
        } else {
          ok = kids[k] < 0;
        }
      }
      if (ok && is_ans[o] != (i == q_ptr[q + 1] - 1)) ok = false;   // exactly the root answers
      if (!ok) { delete sc; return fail(N2NMN_ERR_ARG, "malformed expression node " + std::to_string(i)); }
      r.op = o; r.t = t_idx[i]; r.b = b_idx[i];
      r.in0 = in0[i]; r.in1 = in1[i];
      r.out = is_ans[o] ? q : i;
      r.text = -1;
      r.aux = (o == OP_SCENE) ? scene_bits : -1;
      r.aux2 = -1; r.s0 = r.s1 = r.so = -1;
    }
  }
  if (int rc = finalize_schedule(c->shp, c->cfg.max_batch, &S)) {
    delete sc;
    return fail(rc, "finalize_schedule failed");
  }
  *out = sc;
  return 0;
}

int n2nmn_sched_destroy(n2nmn_sched* s) { delete s; return 0; }

int n2nmn_sched_get_info(const n2nmn_sched* s, n2nmn_sched_info* info) {
  if (!s || !info) return fail(N2NMN_ERR_ARG, "null argument");
  account_schedule(s->shp, const_cast<HostSchedule*>(&s->hs));
  const HostSchedule& S = s->hs;
  info->num_questions = (int)S.q_ptr.size() - 1;
  info->num_valid = S.num_valid;
  info->num_nodes = (int)S.nodes.size();
  info->max_depth = S.max_depth;
  info->num_text_nodes = (int)S.text_t.size();
  info->num_find_nodes = S.num_find_nodes;
  info->num_proj_tiles = (int)S.work.size();
  info->num_launches = (S.groups.empty() ? 0 : 1) + (S.work.empty() ? 0 : 1) + 1 +
                       (S.head_work.empty() ? 0 : 1);
  info->algorithmic_bytes = S.per_node_bytes;
  info->algorithmic_flops = S.per_node_flops;
  for (int k = 0; k < 3; ++k) { info->kernel_bytes[k] = S.kbytes[k]; info->kernel_flops[k] = S.kflops[k]; }
  info->bwd_gemm_flops = S.train ? 2ll * (int64_t)S.entries.size() * s->shp.H * s->shp.W *
                                       s->shp.Dk * s->shp.M : 0;
  return 0;
}

int n2nmn_last_step_info(const n2nmn_ctx* c, n2nmn_sched_info* info) {
  if (!c) return fail(N2NMN_ERR_ARG, "null context");
  return n2nmn_sched_get_info(&c->step_sched, info);
}

int n2nmn_sched_get_nodes(const n2nmn_sched* s, int32_t* out6, int cap) {
  if (!s || !out6) return fail(N2NMN_ERR_ARG, "null argument");
  const HostSchedule& S = s->hs;
  if (cap < (int)S.nodes.size()) return fail(N2NMN_ERR_CAPACITY, "node buffer too small");
  for (size_t i = 0; i < S.nodes.size(); ++i) {
    const NodeRec& r = S.nodes[i];
    int32_t* o = out6 + 6 * i;
    o[0] = r.op; o[1] = r.t; o[2] = r.b; o[3] = S.depth[i]; o[4] = r.in0; o[5] = r.in1;
  }
  return (int)S.nodes.size();
}

int n2nmn_run_schedule(n2nmn_ctx* c, n2nmn_sched* s, float* scores, float* att_arena,
                       void* stream) {
  if (!c || !s || !scores) return fail(N2NMN_ERR_ARG, "null argument");
  if (int rc = check_ready(c)) return rc;
  for (const NodeRec& r : s->hs.nodes)
    if (r.b >= c->N || r.t >= c->T)
      return fail(N2NMN_ERR_ARG, "schedule refers to a batch/time index outside the bound inputs");
  if ((int)s->hs.img_ptr.size() - 1 > c->N && s->hs.img_ptr.back() != s->hs.img_ptr[c->N])
    return fail(N2NMN_ERR_ARG, "schedule image range exceeds the bound inputs");
  float* arena = att_arena ? att_arena : c->arena;
  if (!att_arena && (int)s->hs.nodes.size() > c->arena_slots)
    return fail(N2NMN_ERR_CAPACITY, "too many nodes for the context arena");
  if (s->hs.num_seg != 1 || c->md.num_seg != 1)
    return fail(N2NMN_ERR_STATE, "n2nmn_run_schedule works on a single bound batch");
  return run_tables(c, s, &scores, arena, static_cast<cudaStream_t>(stream), false,
                    att_arena != nullptr);
}

namespace {
int module_fwd_impl(n2nmn_ctx* c, int op, const float* in0, const float* in1,
                    const int32_t* t_idx, const int32_t* b_idx, int n, float* out, void* stream,
                    float scene_val);
}

int n2nmn_module_fwd(n2nmn_ctx* c, int op, const float* in0, const float* in1,
                     const int32_t* t_idx, const int32_t* b_idx, int n, float* out,
                     void* stream) {
  return module_fwd_impl(c, op, in0, in1, t_idx, b_idx, n, out, stream, 3.0f);
}

int n2nmn_scene_fwd(n2nmn_ctx* c, int n, float pos_val, float* out, void* stream) {
  return module_fwd_impl(c, OP_SCENE, nullptr, nullptr, nullptr, nullptr, n, out, stream, pos_val);
}

namespace {
int module_fwd_impl(n2nmn_ctx* c, int op, const float* in0, const float* in1,
                    const int32_t* t_idx, const int32_t* b_idx, int n, float* out, void* stream,
                    float scene_val) {
  if (!c) return fail(N2NMN_ERR_ARG, "null context");
  if (op < 0 || op >= NUM_OPS) return fail(N2NMN_ERR_ARG, "bad opcode");
  if (n == 0) return 0;   // TF Fold's zero-size batches: nothing to do
  if (n < 0 || !out) return fail(N2NMN_ERR_ARG, "bad arguments");
  if (int rc = check_ready(c)) return rc;
  static const int arity[NUM_OPS] = {0, 0, 1, 1, 1, 2, 2, 1, 1, 2, 2, 2, 2, 1};
  static const bool is_ans[NUM_OPS] = {0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1};
  static const bool needs_idx[NUM_OPS] = {0, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 1, 1};
  if ((arity[op] >= 1 && !in0) || (arity[op] >= 2 && !in1))
    return fail(N2NMN_ERR_ARG, "missing attention input");
  if (needs_idx[op] && (!t_idx || !b_idx))
    return fail(N2NMN_ERR_ARG, "time_idx / batch_idx required for this module");
  if (3 * n > c->arena_slots || n > c->cfg.max_batch * c->cfg.max_T)
    return fail(N2NMN_ERR_CAPACITY, "n exceeds the context capacity for a single module call");
  if (c->md.num_seg != 1) return fail(N2NMN_ERR_STATE, "module calls need a single bound batch");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  n2nmn_sched* sc = &c->module_sched;
  sc->uid = g_uid++;   // tables change every call
  HostSchedule& S = sc->hs;
  S.reset();
  S.N = c->N; S.T = c->T;
  S.nodes.resize(n); S.depth.assign(n, 1); S.q_ptr.resize(n + 1);
  int scene_bits;
  std::memcpy(&scene_bits, &scene_val, 4);
  for (int i = 0; i < n; ++i) {
    NodeRec& r = S.nodes[i];
    r.op = op;
    r.t = t_idx ? t_idx[i] : 0;
    r.b = b_idx ? b_idx[i] : 0;
    if (needs_idx[op] && (r.b < 0 || r.b >= c->N || r.t < 0 || r.t >= c->T))
      return fail(N2NMN_ERR_ARG, "time_idx / batch_idx out of range");
    if (!needs_idx[op]) { r.t = 0; r.b = 0; }
    r.in0 = arity[op] >= 1 ? i : -1;
    r.in1 = arity[op] >= 2 ? n + i : -1;
    r.out = is_ans[op] ? i : 2 * n + i;
    r.text = -1;
    r.aux = (op == OP_SCENE) ? scene_bits : -1;
    r.aux2 = -1; r.s0 = r.s1 = r.so = -1;
    S.q_ptr[i] = i;
  }
  S.q_ptr[n] = n;
  if (int rc = finalize_schedule(c->shp, c->N, &S)) return fail(rc, "finalize_schedule failed");
  const size_t map_bytes = (size_t)n * c->HW * sizeof(float);
  if (arity[op] >= 1)
    CUDA_TRY(cudaMemcpyAsync(c->arena, in0, map_bytes, cudaMemcpyDeviceToDevice, st));
  if (arity[op] >= 2)
    CUDA_TRY(cudaMemcpyAsync(c->arena + (size_t)n * c->HW, in1, map_bytes,
                             cudaMemcpyDeviceToDevice, st));
  float* scores = is_ans[op] ? out : c->scores_tmp;
  if (int rc = run_tables(c, sc, &scores, c->arena, st, /*force_wave=*/true)) return rc;
  if (!is_ans[op])
    CUDA_TRY(cudaMemcpyAsync(out, c->arena + (size_t)2 * n * c->HW, map_bytes,
                             cudaMemcpyDeviceToDevice, st));
  return 0;
}
}  // namespace

int n2nmn_forward_group(n2nmn_ctx* c, int num_batches, const float* const* feat_dev,
                        const float* const* wv_dev, const int32_t* const* tokens, int T, int N,
                        const int32_t* vocab_ops, int num_vocab, float* const* scores_dev,
                        uint8_t* const* validity_out, void* stream) {
  if (!c || !feat_dev || !wv_dev || !tokens || !vocab_ops || !scores_dev || num_batches <= 0)
    return fail(N2NMN_ERR_ARG, "null argument");
  for (int i = 0; i < num_batches; ++i)
    if (!tokens[i] || !scores_dev[i]) return fail(N2NMN_ERR_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(c->device));   // callers may drive one context per host thread
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int rc = bind_segments(c, num_batches, feat_dev, wv_dev, N, T, st)) return rc;
  if (int rc = check_ready(c)) return rc;
  n2nmn_sched* sc = &c->step_sched;
  sc->uid = g_uid++;
  const char* err = nullptr;
  const bool direct = !(c->cfg.flags & N2NMN_FLAG_WAVE_EXECUTOR);
  if (int rc = compile_schedule_group(c->shp, tokens, num_batches, T, N, vocab_ops, num_vocab,
                                      &sc->hs, &err, false, direct))
    return fail(rc, err ? err : "compile_schedule failed");
  if (validity_out)
    for (int i = 0; i < num_batches; ++i)
      if (validity_out[i]) std::memcpy(validity_out[i], sc->hs.validity.data() + (size_t)i * N, N);
  if ((int)sc->hs.nodes.size() > c->arena_slots)
    return fail(N2NMN_ERR_CAPACITY, "too many nodes for the context arena");
  return run_tables(c, sc, scores_dev, c->arena, st);
}

int n2nmn_forward_tokens(n2nmn_ctx* c, const float* feat_dev, const float* wv_dev,
                         const int32_t* tokens, int T, int N, const int32_t* vocab_ops,
                         int num_vocab, float* scores_dev, uint8_t* validity_out, void* stream) {
  if (!c || !tokens || !vocab_ops || !scores_dev) return fail(N2NMN_ERR_ARG, "null argument");
  return n2nmn_forward_group(c, 1, &feat_dev, &wv_dev, &tokens, T, N, vocab_ops, num_vocab,
                             &scores_dev, validity_out ? &validity_out : nullptr, stream);
}

namespace {
// Host-buffer variant of n2nmn_forward_group: H2D of every batch's features and word vectors into
// context-owned staging, the kernels, D2H of every batch's scores — all enqueued on `stream`.
int forward_host_impl(n2nmn_ctx* c, int nb, const void* const* feat_host,
                      const float* const* wv_host, const int32_t* const* tokens, int T, int N,
                      const int32_t* vocab_ops, int num_vocab, float* const* scores_host,
                      uint8_t* const* validity_out, void* stream, bool sync, bool feat_f16 = false) {
  if (!c || !feat_host || !wv_host || !tokens || !scores_host || nb <= 0)
    return fail(N2NMN_ERR_ARG, "null argument");
  if (nb > c->G) return fail(N2NMN_ERR_CAPACITY, "too many batches for one group");
  if (N <= 0 || N > c->cfg.max_batch || T <= 0 || T > c->cfg.max_T)
    return fail(N2NMN_ERR_CAPACITY, "N or T exceeds the context capacity");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t fcap = (size_t)c->cfg.max_batch * c->HW * c->cfg.D;
  const size_t wcap = (size_t)c->cfg.max_T * c->cfg.max_batch * c->cfg.text_dim;
  const size_t scap = (size_t)c->cfg.max_batch * c->cfg.num_choices;
  const size_t fbytes = (size_t)N * c->HW * c->cfg.D * sizeof(float);
  const size_t wbytes = (size_t)T * N * c->cfg.text_dim * sizeof(float);
  const size_t sbytes = (size_t)N * c->cfg.num_choices * sizeof(float);
  if (!c->e2e_feat) {
    CUDA_TRY(cudaMalloc(&c->e2e_feat, c->G * fcap * sizeof(float)));
    CUDA_TRY(cudaMalloc(&c->e2e_wv, c->G * wcap * sizeof(float)));
    CUDA_TRY(cudaMalloc(&c->e2e_scores, c->G * scap * sizeof(float)));
  }
  if (feat_f16 && (fbytes / sizeof(float)) % 8 != 0)
    return fail(N2NMN_ERR_ARG, "fp16 host features need N*H*W*D to be a multiple of 8");
  const size_t fcap16 = (fcap + 7) & ~(size_t)7;   // fp16 staging slots stay 16-byte aligned
  if (feat_f16 && !c->e2e_feat_f16)
    CUDA_TRY(cudaMalloc(&c->e2e_feat_f16, c->G * fcap16 * sizeof(uint16_t)));
  const float* fd[kMaxSeg];
  const float* wd[kMaxSeg];
  float* sd[kMaxSeg];
  for (int i = 0; i < nb; ++i) {
    if (!feat_host[i] || !wv_host[i] || !scores_host[i]) return fail(N2NMN_ERR_ARG, "null argument");
    fd[i] = c->e2e_feat + i * fcap; wd[i] = c->e2e_wv + i * wcap; sd[i] = c->e2e_scores + i * scap;
    if (feat_f16)
      CUDA_TRY(cudaMemcpyAsync(c->e2e_feat_f16 + i * fcap16, feat_host[i], fbytes / 2,
                               cudaMemcpyHostToDevice, st));
    else
      CUDA_TRY(cudaMemcpyAsync(c->e2e_feat + i * fcap, feat_host[i], fbytes, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(c->e2e_wv + i * wcap, wv_host[i], wbytes, cudaMemcpyHostToDevice, st));
  }
  if (feat_f16) {   // widen every staged grid to the fp32 layout the kernels read
    const size_t cnt = fbytes / sizeof(float);
    for (int i = 0; i < nb; ++i) {
      const size_t n8 = (cnt + 7) / 8;
      widen_f16_kernel<<<(unsigned)std::min<size_t>((n8 + 255) / 256, 148 * 8), 256, 0, st>>>(
          reinterpret_cast<const uint4*>(c->e2e_feat_f16 + i * fcap16),
          reinterpret_cast<float4*>(c->e2e_feat + i * fcap), n8);
      ++c->launches;
    }
    CUDA_TRY(cudaGetLastError());
  }
  int rc = n2nmn_forward_group(c, nb, fd, wd, tokens, T, N, vocab_ops, num_vocab, sd, validity_out,
                               stream);
  if (rc == 0) {
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < nb && e == cudaSuccess; ++i)
      e = cudaMemcpyAsync(scores_host[i], sd[i], sbytes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && sync) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fail(N2NMN_ERR_CUDA, cudaGetErrorString(e));
  } else if (sync) {
    cudaStreamSynchronize(st);
  }
  return rc;
}
}  // namespace

int n2nmn_forward_host(n2nmn_ctx* c, const float* feat_host, const float* wv_host,
                       const int32_t* tokens, int T, int N, const int32_t* vocab_ops,
                       int num_vocab, float* scores_host, uint8_t* validity_out, void* stream) {
  return forward_host_impl(c, 1, reinterpret_cast<const void* const*>(&feat_host), &wv_host, &tokens, T, N, vocab_ops, num_vocab,
                           &scores_host, validity_out ? &validity_out : nullptr, stream, true);
}

int n2nmn_forward_host_async(n2nmn_ctx* c, const float* feat_host, const float* wv_host,
                             const int32_t* tokens, int T, int N, const int32_t* vocab_ops,
                             int num_vocab, float* scores_host, uint8_t* validity_out,
                             void* stream) {
  return forward_host_impl(c, 1, reinterpret_cast<const void* const*>(&feat_host), &wv_host, &tokens, T, N, vocab_ops, num_vocab,
                           &scores_host, validity_out ? &validity_out : nullptr, stream, false);
}

int n2nmn_forward_group_host_async(n2nmn_ctx* c, int num_batches, const float* const* feat_host,
                                   const float* const* wv_host, const int32_t* const* tokens,
                                   int T, int N, const int32_t* vocab_ops, int num_vocab,
                                   float* const* scores_host, uint8_t* const* validity_out,
                                   void* stream) {
  return forward_host_impl(c, num_batches, reinterpret_cast<const void* const*>(feat_host), wv_host,
                           tokens, T, N, vocab_ops, num_vocab, scores_host, validity_out, stream,
                           false);
}

int n2nmn_forward_group_host_f16_async(n2nmn_ctx* c, int num_batches,
                                       const uint16_t* const* feat_host_f16,
                                       const float* const* wv_host, const int32_t* const* tokens,
                                       int T, int N, const int32_t* vocab_ops, int num_vocab,
                                       float* const* scores_host, uint8_t* const* validity_out,
                                       void* stream) {
  return forward_host_impl(c, num_batches, reinterpret_cast<const void* const*>(feat_host_f16),
                           wv_host, tokens, T, N, vocab_ops, num_vocab, scores_host, validity_out,
                           stream, false, true);
}

int n2nmn_max_group(const n2nmn_ctx* c) { return c ? c->G : 0; }


int64_t n2nmn_flat_size(const n2nmn_ctx* c) { return c ? c->flat_size : 0; }

int n2nmn_flat_offset(const n2nmn_ctx* c, int index, int64_t* offset, int64_t* count) {
  if (!c || index < 0 || index >= (int)c->vars.size()) return fail(N2NMN_ERR_ARG, "bad index");
  if (offset) *offset = c->flat_offset[index];
  if (count) *count = (int64_t)c->vars[index].count;
  return 0;
}

namespace {
int ensure_repack_tables(n2nmn_ctx* c) {
  const int nv = (int)c->vars.size();
  if (!c->d_repack) {   // what n2nmn_set_weight does per variable, as device tables
    std::vector<RepackSeg> segs(nv);
    ProjRepack& pr = c->proj_repack;
    for (int s = 0; s < NUM_PROJ_SETS; ++s) { pr.w_off[s] = pr.b_off[s] = -1; pr.wt[s] = pr.bias[s] = nullptr; }
    c->proj_repack_sets = 0;
    for (int i = 0; i < nv; ++i) {
      const Variable& v = c->vars[i];
      segs[i].src_off = (int)c->flat_offset[i];
      segs[i].count = (int)v.count;
      segs[i].cols = (int)v.shape.back();
      segs[i].kind = v.kind == VK_PITCHED ? 1 : 0;
      segs[i].dst_off = (long long)v.offset;
      if (v.kind == VK_PROJ_W) {
        pr.w_off[v.set] = (int)c->flat_offset[i];
        pr.wt[v.set] = c->proj_wt[v.set];
        pr.bias[v.set] = c->proj_bias[v.set];
        pr.set_of_z[c->proj_repack_sets++] = v.set;
      } else if (v.kind == VK_PROJ_B) {
        pr.b_off[v.set] = (int)c->flat_offset[i];
      }
    }
    CUDA_TRY(cudaMalloc(&c->d_repack, nv * sizeof(RepackSeg)));
    CUDA_TRY(cudaMemcpy(c->d_repack, segs.data(), nv * sizeof(RepackSeg), cudaMemcpyHostToDevice));
  }
  return 0;
}

// the derived copies: K-major padded projection weights / biases and the Transform quadratic form
int repack_derived(n2nmn_ctx* c, const float* wflat_dev, cudaStream_t st) {
  for (size_t i = 0; i < c->vars.size(); ++i)
    for (int os = 0; os < NUM_OUT_SETS; ++os)
      if (c->vars[i].slot == &c->md.out_w[os] && c->out_wp[os]) {
        pitch_rows_kernel<<<(unsigned)c->cfg.map_dim, 256, 0, st>>>(
            wflat_dev + c->flat_offset[i], c->cfg.map_dim, c->cfg.num_choices, c->out_wp[os], c->Cp);
        if (c->out_wt_hi[os])
          out_wt_split_kernel<<<dim3((c->Cpad + 31) / 32, (c->Mp + 31) / 32), dim3(32, 8), 0, st>>>(
              wflat_dev + c->flat_offset[i], c->cfg.map_dim, c->cfg.num_choices, c->out_wt_hi[os],
              c->out_wt_lo[os], c->Mp, c->Cpad);
        ++c->launches;
      }
  if (c->proj_repack_sets > 0) {
    dim3 grid((c->Kp + 31) / 32, (c->Mp + 31) / 32, c->proj_repack_sets), block(32, 8);
    proj_repack_kernel<<<grid, block, 0, st>>>(wflat_dev, c->proj_repack, c->Dk, c->cfg.map_dim,
                                               c->Kp, c->Mp);
    ++c->launches;
  }
  if (c->conv_quad) {
    conv_quad_kernel<<<quad_pitch(c->cfg.kernel_size), 256, 0, st>>>(
        c->md.conv_k, c->md.conv_b, c->md.elt_w[ES_TRANSFORM], c->cfg.kernel_size, c->cfg.map_dim,
        c->Mp, c->conv_quad);
    ++c->launches;
  }
  CUDA_TRY(cudaGetLastError());
  for (Variable& v : c->vars) v.loaded = true;
  return 0;
}
}  // namespace

int n2nmn_load_flat_weights(n2nmn_ctx* c, const float* wflat_dev, void* stream) {
  if (!c || !wflat_dev) return fail(N2NMN_ERR_ARG, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int rc = ensure_repack_tables(c)) return rc;
  repack_all_kernel<<<dim3(16, (unsigned)c->vars.size()), 256, 0, st>>>(wflat_dev, c->d_repack,
                                                                        c->wbuf, c->Mp);
  ++c->launches;
  return repack_derived(c, wflat_dev, st);
}

int n2nmn_train_backward(n2nmn_ctx* c, const float* feat_dev, const float* wv_dev,
                         const int32_t* tokens, int T, int N, const int32_t* vocab_ops,
                         int num_vocab, const int32_t* labels_host, float invalid_expr_loss,
                         float* scores_dev, float* gflat_dev, float* dword_dev, float* loss_dev,
                         uint8_t* validity_out, void* stream) {
  if (!c || !tokens || !vocab_ops || !labels_host || !scores_dev || !gflat_dev || !loss_dev)
    return fail(N2NMN_ERR_ARG, "null argument");
  if (c->cfg.flags & N2NMN_FLAG_WAVE_EXECUTOR)
    return fail(N2NMN_ERR_ARG, "training uses the tree executor");
  // The backward walk is instantiated for the 3x3 / 5x5 Transform families and sizes its channel
  // loops for Mp <= 512 (backward.cuh); the VQA family (no conv Transform, Mp = 1024) has no
  // backward path yet.
  if (c->cfg.family == N2NMN_VQA || c->Mp > 512)
    return fail(N2NMN_ERR_ARG, "n2nmn_train_backward: the VQA family / map_dim > 512 is not supported");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int rc = n2nmn_bind_inputs(c, feat_dev, wv_dev, N, T, stream)) return rc;
  if (int rc = check_ready(c)) return rc;
  const int NB = c->cfg.max_batch, TT = c->cfg.max_T, C = c->cfg.num_choices;
  if (!c->dscores) {
    CUDA_TRY(cudaMalloc(&c->dscores, (size_t)NB * C * sizeof(float)));
    CUDA_TRY(cudaMalloc(&c->per_sample, (size_t)NB * sizeof(float)));
    CUDA_TRY(cudaMalloc(&c->dtau, (size_t)c->text_rows_cap * c->Mp * sizeof(float)));
    c->dmap_entries = NB * TT;
    CUDA_TRY(cudaMalloc(&c->dmap, (size_t)c->dmap_entries * c->HW * c->Mp * sizeof(float)));
    CUDA_TRY(cudaMalloc(&c->dstencil, (size_t)c->dmap_entries * c->HW * c->Mp * sizeof(float)));
    CUDA_TRY(cudaMalloc(&c->gmap, (size_t)c->arena_slots * ((c->HW + 3) & ~3) * sizeof(float)));
    CUDA_TRY(cudaMalloc(&c->phi_buf, (size_t)NB * 2 * c->Mp * sizeof(float)));
    c->wg_ok = c->Mp == kWgN && c->Dk % kWgM == 0 && !c->feat_aug;
    if (c->wg_ok) {
      if (int rc = encode_mn_blocks(c, &c->wg_maps.b, c->dmap, c->Mp, c->HW, c->dmap_entries, c->Mp,
                                    kWgP, kWgN / 32))
        return rc;
      CUDA_TRY(cudaFuncSetAttribute(wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)kWgSmemBytes));
    }
    const BwdSmem L = bwd_smem_layout(c->cfg.H, c->cfg.W, c->Mp, c->cfg.kernel_size, C);
    const int bwd_smem = (int)(L.total * sizeof(float));
    CUDA_TRY(cudaFuncSetAttribute(tree_bwd_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem));
    CUDA_TRY(cudaFuncSetAttribute(tree_bwd_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem));
    CUDA_TRY(cudaFuncSetAttribute(tree_bwd_kernel<5, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem));
    CUDA_TRY(cudaFuncSetAttribute(tree_bwd_kernel<5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem));
    CUDA_TRY(cudaFuncSetAttribute(tree_bwd_kernel<3, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CUDA_TRY(cudaFuncSetAttribute(tree_bwd_kernel<5, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CUDA_TRY(cudaFuncSetAttribute(xtb_mma_kernel<FeatGradSrc>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kXtbSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(xtb_mma_kernel<TextGradSrc>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kXtbSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(text_xgrad_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kXgSmemBytes));
  }
  n2nmn_sched* sc = &c->step_sched;
  sc->uid = g_uid++;
  const char* err = nullptr;
  if (int rc = compile_schedule(c->shp, tokens, T, N, vocab_ops, num_vocab, &sc->hs, &err, true))
    return fail(rc, err ? err : "compile_schedule failed");
  {   // every node bucketed by depth (leaves = 1): the backward runs one launch per level, top down
    HostSchedule& W = sc->hs;
    // bucket 2*depth: the level's Transform nodes (the long CTAs: scheduled first), 2*depth + 1:
    // its other nodes
    const int nb = 2 * (W.max_depth + 1);
    W.bwd_ptr.assign(nb + 1, 0);
    auto bucket = [&](size_t i) { return 2 * W.depth[i] + (W.nodes[i].op == OP_TRANSFORM ? 0 : 1); };
    for (size_t i = 0; i < W.nodes.size(); ++i) ++W.bwd_ptr[bucket(i) + 1];
    for (int b = 0; b < nb; ++b) W.bwd_ptr[b + 1] += W.bwd_ptr[b];
    W.bwd_nodes.assign(W.nodes.size(), 0);
    std::vector<int32_t> fill(W.bwd_ptr.begin(), W.bwd_ptr.end() - 1);
    for (size_t i = 0; i < W.nodes.size(); ++i) W.bwd_nodes[fill[bucket(i)]++] = (int32_t)i;
    W.entry_order.resize(W.entries.size());
    for (size_t i = 0; i < W.entries.size(); ++i) W.entry_order[i] = (int32_t)i;
    std::stable_sort(W.entry_order.begin(), W.entry_order.end(),
                     [&](int32_t a, int32_t b) { return W.entries[a].set < W.entries[b].set; });
  }
  const HostSchedule& S = sc->hs;
  if (validity_out) std::memcpy(validity_out, S.validity.data(), N);
  if ((int)S.nodes.size() > c->arena_slots || (int)S.entries.size() > c->dmap_entries ||
      (int)S.nodes.size() > c->dmap_entries)
    return fail(N2NMN_ERR_CAPACITY, "too many nodes for the context");
  if (S.max_stack > c->stack_cap)
    return fail(N2NMN_ERR_CAPACITY, "layout too deep for the training path");
  for (int i = 0; i < N; ++i)
    if (labels_host[i] < 0 || labels_host[i] >= C) return fail(N2NMN_ERR_ARG, "label out of range");
  // ---- forward (keeps every attention map in the context arena + the stored maps)
  c->train_labels = labels_host;
  const int rc = run_tables(c, sc, &scores_dev, c->arena, st, false, /*write_arena=*/true);
  c->train_labels = nullptr;
  if (rc) return rc;
  const uint8_t* d = c->last_tables;
  const TableOffsets& o = c->last_offsets;
  const NodeRec* d_nodes = reinterpret_cast<const NodeRec*>(d + o.nodes);
  const int32_t* d_qptr = reinterpret_cast<const int32_t*>(d + o.q_ptr);
  // ---- loss and d(scores)
  CUDA_TRY(cudaMemsetAsync(gflat_dev, 0, (size_t)c->flat_size * sizeof(float), st));
  CUDA_TRY(cudaMemsetAsync(loss_dev, 0, sizeof(float), st));
  if (dword_dev)
    CUDA_TRY(cudaMemsetAsync(dword_dev, 0, (size_t)T * N * c->cfg.text_dim * sizeof(float), st));
  loss_kernel<<<(N + 7) / 8, 256, 0, st>>>(scores_dev, reinterpret_cast<const int32_t*>(d + o.labels),
                                           d_qptr, N, C, invalid_expr_loss, c->dscores,
                                           loss_dev + 1, loss_dev);
  ++c->launches;
  prof_mark(c, "loss_kernel", st);
  // ---- reverse tree walk
  BwdCtx bc;
  bc.md = c->md; bc.tb = c->tb; bc.arena = c->arena; bc.scores = scores_dev;
  bc.dscores = c->dscores; bc.mbuf = c->mbuf; bc.gflat = gflat_dev; bc.dtau = c->dtau;
  bc.dmap = c->dmap; bc.dstencil = c->dstencil; bc.gmap = c->gmap; bc.phi = c->phi_buf; bc.go = c->go;
  const BwdSmem L = bwd_smem_layout(c->cfg.H, c->cfg.W, c->Mp, c->cfg.kernel_size, C);
  const size_t bsm = L.total * sizeof(float);
  const int32_t* d_entry = reinterpret_cast<const int32_t*>(d + o.node_entry);
  const int32_t* d_bwd = reinterpret_cast<const int32_t*>(d + o.bwd_nodes);
  CUDA_TRY(cudaMemsetAsync(c->gmap, 0, S.nodes.size() * (size_t)L.HWp * sizeof(float), st));
  CUDA_TRY(cudaMemsetAsync(c->dtau, 0, S.text_t.size() * (size_t)c->Mp * sizeof(float), st));
  // a level's nodes are listed [Transform nodes | others]; a level with Transform nodes runs the
  // full instantiation over all of them (one CTA per SM), a level without the lighter one
  for (int dd = S.max_depth; dd >= 1; --dd) {
    const int first = S.bwd_ptr[2 * dd], cnt = S.bwd_ptr[2 * dd + 2] - first;
    if (cnt <= 0) continue;
    const int n_tr = S.bwd_ptr[2 * dd + 1] - S.bwd_ptr[2 * dd];
    const bool has_tr = n_tr > 0;
    // CTAs per splittable node: as many as keep the level's heavy CTAs within one wave of the
    // SMs (148 at one CTA per SM with Transform nodes, 296 without)
    const int slices = has_tr ? std::max(3, std::min(kBwdSlicesMax, 148 / n_tr))
                              : std::max(2, std::min(6, 296 / cnt));
    cudaLaunchConfig_t bl;
    std::memset(&bl, 0, sizeof(bl));
    bl.gridDim = dim3(cnt, slices);
    bl.blockDim = dim3(kNodeThreads);
    bl.dynamicSmemBytes = bsm;
    bl.stream = st;
    cudaLaunchAttribute battr[1];
    battr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    battr[0].val.programmaticStreamSerializationAllowed = 1;
    bl.attrs = battr;
    bl.numAttrs = c->use_pdl ? 1 : 0;
    const bool k5 = c->cfg.kernel_size == 5;
    if (has_tr) {
      if (k5) CUDA_TRY(cudaLaunchKernelEx(&bl, tree_bwd_kernel<5, true>, bc, d_nodes, d_bwd, first, d_entry));
      else CUDA_TRY(cudaLaunchKernelEx(&bl, tree_bwd_kernel<3, true>, bc, d_nodes, d_bwd, first, d_entry));
    } else {
      if (k5) CUDA_TRY(cudaLaunchKernelEx(&bl, tree_bwd_kernel<5, false>, bc, d_nodes, d_bwd, first, d_entry));
      else CUDA_TRY(cudaLaunchKernelEx(&bl, tree_bwd_kernel<3, false>, bc, d_nodes, d_bwd, first, d_entry));
    }
    ++c->launches;
  }
  prof_mark(c, "tree_bwd_kernel", st);
  // ---- text layers
  const int rows = (int)S.text_t.size();
  if (rows > 0) {
    const int32_t* d_tt = reinterpret_cast<const int32_t*>(d + o.text_t);
    const int32_t* d_tb = reinterpret_cast<const int32_t*>(d + o.text_b);
    const int32_t* d_ss = reinterpret_cast<const int32_t*>(d + o.text_set_start);
    const bool exact = (c->cfg.flags & N2NMN_FLAG_PROJ_FP32_SIMT) != 0;
    if (exact) {
      dim3 g1((c->cfg.text_dim + 7) / 8, NUM_TEXT_SETS);
      text_wgrad_kernel<<<g1, 256, 0, st>>>(c->md, c->dtau, d_tt, d_tb, d_ss, gflat_dev, c->go);
    } else {
      TextGradSrc ts{c->md, c->dtau, d_tt, d_tb, d_ss, gflat_dev, c->go};
      dim3 g1((c->cfg.text_dim + kXtbM - 1) / kXtbM, (c->cfg.map_dim + kXtbN - 1) / kXtbN,
              NUM_TEXT_SETS);
      xtb_mma_kernel<TextGradSrc><<<g1, kXtbThreads, kXtbSmemBytes, st>>>(ts, 1);
    }
    ++c->launches;
    if (dword_dev) {
      if (exact) {
        text_xgrad_kernel<<<rows, 256, c->Mp * sizeof(float), st>>>(c->md, c->dtau, d_tt, d_tb, d_ss,
                                                                   dword_dev, c->dword_scale);
      } else {
        TextSetRows tsr;
        int groups = 0;
        for (int i = 0; i <= NUM_TEXT_SETS; ++i) tsr.start[i] = S.text_set_start[i];
        for (int i = 0; i < NUM_TEXT_SETS; ++i) groups += (tsr.start[i + 1] - tsr.start[i] + 63) / 64;
        dim3 g3((c->cfg.text_dim + 63) / 64, groups);
        text_xgrad_mma_kernel<<<g3, kXgThreads, kXgSmemBytes, st>>>(c->md, c->dtau, d_tt, d_tb, tsr,
                                                                    dword_dev, c->dword_scale);
      }
      ++c->launches;
    }
    prof_mark(c, "text_grad_kernels", st);
  }
  // ---- feature-side layers: dW_set = Σ X^T·B
  const int ne = (int)S.entries.size();
  if (ne > 0) {
    const BwdEntry* d_ent = reinterpret_cast<const BwdEntry*>(d + o.entries);
    if (c->cfg.flags & N2NMN_FLAG_PROJ_FP32_SIMT) {
      const int chunks = std::min(ne, 16);
      const int per = (ne + chunks - 1) / chunks;
      dim3 g2((c->Dk + kFgTile - 1) / kFgTile, (c->cfg.map_dim + kFgTile - 1) / kFgTile,
              (ne + per - 1) / per);
      feat_grad_kernel<<<g2, 256, 0, st>>>(c->md, c->dmap, d_ent, ne, per, gflat_dev, c->go);
    } else if (c->wg_ok && !std::getenv("N2NMN_WGRAD_MMA_SYNC")) {
      // tcgen05, both operands MN-major straight from the feature grid and the B maps
      if (int rc = encode_mn_blocks(c, &c->wg_maps.x, c->md.feat, c->Dk, c->HW, N, c->md.feat_pitch,
                                    kWgP, kWgM / 32))
        return rc;
      const int slabs = c->Dk / kWgM;
      const int chunks = std::max(1, std::min(ne, 148 / slabs));
      WgradParams wp;
      wp.entries = d_ent;
      wp.order = reinterpret_cast<const int32_t*>(d + o.entry_order);
      wp.num_entries = ne; wp.per_cta = (ne + chunks - 1) / chunks;
      wp.HW = c->HW; wp.Dk = c->Dk; wp.M = c->cfg.map_dim; wp.gflat = gflat_dev; wp.go = c->go;
      dim3 gw(slabs, (ne + wp.per_cta - 1) / wp.per_cta);
      wgrad_umma_kernel<<<gw, kWgThreads, kWgSmemBytes, st>>>(c->wg_maps, wp);
      prof_mark(c, "feat_grad_kernel", st);
      bmap_colsum_kernel<<<ne, 1024, 0, st>>>(c->dmap, d_ent, c->HW, c->cfg.map_dim, c->Mp, gflat_dev,
                                             c->go);
      ++c->launches;
      ++c->launches;
      prof_mark(c, "bias_grad_kernel", st);
      CUDA_TRY(cudaGetLastError());
      return 0;
    } else {
      // two CTAs per SM (106 KB of ring each): ~296 CTAs over (Dk/128) x (M/64) tiles
      const int tiles = ((c->Dk + kXtbM - 1) / kXtbM) * ((c->cfg.map_dim + kXtbN - 1) / kXtbN);
      const int chunks = std::max(1, std::min(ne, (2 * 148 + tiles - 1) / tiles));
      const int per = (ne + chunks - 1) / chunks;
      FeatGradSrc fs{c->md, c->dmap, d_ent, ne, gflat_dev, c->go};
      dim3 g2((c->Dk + kXtbM - 1) / kXtbM, (c->cfg.map_dim + kXtbN - 1) / kXtbN,
              (ne + per - 1) / per);
      xtb_mma_kernel<FeatGradSrc><<<g2, kXtbThreads, kXtbSmemBytes, st>>>(fs, per);
    }
    ++c->launches;
    prof_mark(c, "feat_grad_kernel", st);
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

namespace {
int adam_impl(n2nmn_ctx* c, float* wflat, float* gflat, float* m, float* v, int step, float lr,
              float beta1, float beta2, float eps, float max_norm, float weight_decay, float gscale,
              float* l2_dev, cudaStream_t st) {
  const int nv = (int)c->vars.size();
  if (!c->d_segs) {
    std::vector<VarSeg> segs(nv);
    for (int i = 0; i < nv; ++i) {
      segs[i].offset = (int)c->flat_offset[i];
      segs[i].count = (int)c->vars[i].count;
      const std::string& n = c->vars[i].name;   // l2_reg covers ".../weights" only
      segs[i].decay = n.size() >= 8 && n.compare(n.size() - 8, 8, "/weights") == 0;
    }
    CUDA_TRY(cudaMalloc(&c->d_segs, nv * sizeof(VarSeg)));
    CUDA_TRY(cudaMemcpy(c->d_segs, segs.data(), nv * sizeof(VarSeg), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&c->d_sumsq, nv * sizeof(float)));
  }
  CUDA_TRY(cudaMemsetAsync(c->d_sumsq, 0, nv * sizeof(float), st));
  dim3 grid(32, nv);
  grad_norm_kernel<<<grid, 256, 0, st>>>(wflat, gflat, c->d_segs, weight_decay, gscale, c->d_sumsq,
                                         l2_dev);
  const double lr_t = (double)lr * std::sqrt(1.0 - std::pow((double)beta2, step)) /
                      (1.0 - std::pow((double)beta1, step));
  if (int rc = ensure_repack_tables(c)) return rc;
  adam_clip_kernel<<<grid, 256, 0, st>>>(wflat, gflat, m, v, c->d_segs, c->d_sumsq, (float)lr_t,
                                         beta1, beta2, eps, max_norm, c->d_repack, c->wbuf, c->Mp);
  c->launches += 2;
  CUDA_TRY(cudaGetLastError());
  prof_mark(c, "clip_adam_kernels", st);
  const int rc = repack_derived(c, wflat, st);
  prof_mark(c, "repack_kernels", st);
  return rc;
}
}  // namespace

int n2nmn_adam_step(n2nmn_ctx* c, float* wflat, float* gflat, float* m, float* v, int step,
                    float lr, float beta1, float beta2, float eps, float max_norm,
                    float weight_decay, void* stream) {
  if (!c || !wflat || !gflat || !m || !v || step < 1) return fail(N2NMN_ERR_ARG, "bad argument");
  return adam_impl(c, wflat, gflat, m, v, step, lr, beta1, beta2, eps, max_norm, weight_decay, 1.f,
                   nullptr, static_cast<cudaStream_t>(stream));
}

int n2nmn_train_finish(n2nmn_ctx* c, float* wflat, float* gflat, float* m, float* v, int step,
                       float lr, float beta1, float beta2, float eps, float max_norm,
                       float weight_decay, const float* loss_sum_dev, const float* per_sample_dev,
                       const float* log_seq_prob_dev, int N, int world, float baseline_decay,
                       const float* state_in_dev, float* state_out_dev, float* coeff_dev,
                       void* stream) {
  if (!c || !wflat || !gflat || !m || !v || step < 1 || !loss_sum_dev || !per_sample_dev ||
      !state_in_dev || !state_out_dev || N <= 0 || world <= 0)
    return fail(N2NMN_ERR_ARG, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  train_scalars_kernel<<<1, 256, 0, st>>>(loss_sum_dev, per_sample_dev, log_seq_prob_dev, N, world,
                                          baseline_decay, state_in_dev, state_out_dev, coeff_dev);
  ++c->launches;
  return adam_impl(c, wflat, gflat, m, v, step, lr, beta1, beta2, eps, max_norm, weight_decay,
                   1.f / (float)world, state_out_dev + 3, st);
}

int n2nmn_set_grad_scale(n2nmn_ctx* c, float scale) {
  if (!c) return fail(N2NMN_ERR_ARG, "null context");
  c->dword_scale = scale;
  return 0;
}

#if defined(N2NMN_EXP_TIMELINE)
extern "C" int n2nmn_exp_set_timeline(long long* dev_buf) {
  return cudaMemcpyToSymbol(n2nmn::g_timeline, &dev_buf, sizeof(dev_buf)) == cudaSuccess ? 0 : -1;
}
#endif

int n2nmn_set_tree_cluster(n2nmn_ctx* c, int ctas_per_question) {
  if (!c) return fail(N2NMN_ERR_ARG, "n2nmn_set_tree_cluster: null context");
  const int v = ctas_per_question;
  if (!(v == 0 || v == 1 || v == 2 || v == 4 || v == 8))
    return fail(N2NMN_ERR_ARG, "n2nmn_set_tree_cluster: ctas_per_question must be 0, 1, 2, 4 or 8");
  c->tree_cluster = v;
  return 0;
}

int n2nmn_set_proj_ctas(n2nmn_ctx* c, int max_ctas) {
  if (!c) return fail(N2NMN_ERR_ARG, "n2nmn_set_proj_ctas: null context");
  if (max_ctas < 0) return fail(N2NMN_ERR_ARG, "n2nmn_set_proj_ctas: max_ctas must be >= 0");
  c->proj_max_ctas = max_ctas;
  return 0;
}

int n2nmn_set_text_ctas_per_group(n2nmn_ctx* c, int n) {
  if (!c) return fail(N2NMN_ERR_ARG, "n2nmn_set_text_ctas_per_group: null context");
  if (n < 0) return fail(N2NMN_ERR_ARG, "n2nmn_set_text_ctas_per_group: n must be >= 0");
  c->text_ctas_per_group = n;
  return 0;
}

int n2nmn_set_profiling(n2nmn_ctx* c, int enabled) {
  if (!c) return fail(N2NMN_ERR_ARG, "null context");
  c->profiling = enabled != 0;
  return 0;
}

int n2nmn_get_launch_times(n2nmn_ctx* c, const char** names, float* us, int cap) {
  if (!c) return fail(N2NMN_ERR_ARG, "null context");
  if (c->ev_used < 2) return 0;
  if (cudaEventSynchronize(c->ev[c->ev_used - 1]) != cudaSuccess)
    return fail(N2NMN_ERR_CUDA, "event sync failed");
  int n = 0;
  for (int i = 1; i < c->ev_used && n < cap; ++i, ++n) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev[i - 1], c->ev[i]);
    if (names) names[n] = c->ev_names[i];
    if (us) us[n] = ms * 1000.f;
  }
  return n;
}

int64_t n2nmn_launch_count(const n2nmn_ctx* c) { return c ? c->launches : 0; }

}  // extern "C"
