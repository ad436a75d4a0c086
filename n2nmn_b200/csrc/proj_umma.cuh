// K2: conv_image contraction on the 5th-gen tensor cores (tcgen05, kind::tf32, cta_group::2) with
// the module epilogue fused on the TMEM accumulator. See proj_common.cuh for the math and what is
// fused.
//
// Tiling. The kernel runs as CTA PAIRS (clusters of 2 = one TPC). One work item = two 128-row
// tiles of the flattened (image, pixel) axis that use the same weight set, one per CTA, x all Mp
// output columns, walked as Mp/256 N-tiles: ONE tcgen05.mma.cta_group::2 of shape 256x256x8 (fp32
// operands read as TF32 straight from the caller's fp32 feature grid — no conversion pass) covers
// both tiles. Each CTA stages its own A slice (128 rows x 32 floats, 16 KB) and only HALF of the
// B slice (128 of the 256 weight columns x 32 floats, 16 KB): the tensor cores of both SMs read
// both halves, so the L2->SM operand traffic per MMA is 32 KB instead of the 48 KB of the
// single-CTA form (the mainloop was bound by exactly that traffic, 4.2x the
// algorithmic bytes). K is streamed in 32-float (128-byte, one swizzle atom) slices through a
// 4-stage TMA ring (SWIZZLE_128B, K-major); the bytes of both CTAs are counted on the LEADER's
// `full` barrier, the MMA completion is multicast to both CTAs' `empty` barriers.
// TMEM of each CTA holds two 128x256 fp32 accumulators (all 512 columns), so the epilogue of
// N-tile i overlaps the MMAs of N-tile i+1 / of the next work item. The grid is persistent: pair i
// walks work items i, i+pairs, ... One launch may span several batches (segments, common.cuh): a
// work item names the segment of each of its tiles and the feature tensor map is picked by it.
//
// Warp roles (384 threads per CTA): warp 0 = TMA producer (one elected lane), warp 1 = TMEM
// allocator + (leader CTA only) MMA issuer, warps 2-3 idle (registers handed to the epilogue with
// setmaxnreg), warps 4..11 = epilogue; epilogue warp w reads TMEM lanes [32*(w%4), 32*(w%4)+32)
// i.e. tile rows with that offset, and the two warps that share a lane quarter split the 256
// columns of an N-tile in halves. Their partial sums meet in shared memory. Each CTA's epilogue
// works on its own tile exactly as in the single-CTA form; it releases the accumulator to the
// leader's MMA warp with a remote mbarrier arrive. DESIGN.md §4 has the anatomy and what was
// measured on the way.
//
// Precision: TF32 operands (10-bit mantissa), fp32 accumulate. Error budget vs the fp32/fp64
// oracle is in DESIGN.md; tests/test_gpu_parity.py holds every attention map to 1e-3 abs.
#pragma once
#include "proj_common.cuh"
#include "ptx_sm100.cuh"

namespace n2nmn {

constexpr int kBM = 128;           // rows per tile = rows per CTA (UMMA M = 256 over the pair)
constexpr int kBN = 256;           // columns per N-tile (UMMA N)
constexpr int kBNHalf = kBN / 2;   // weight columns staged by each CTA
constexpr int kBK = 32;            // fp32 elements per K slice = 128 bytes
constexpr int kUmmaK = 8;          // K per tcgen05.mma for tf32 (32 bytes)
constexpr int kABytes = kBM * kBK * 4;        // 16384
constexpr int kBHalfBytes = kBNHalf * kBK * 4;   // 16384
constexpr int kStageBytes = kABytes + kBHalfBytes;
// Ring depth. The MMAs of a slice take ~512 cycles per SM, a TMA box under load needs 1.5-2 K
// cycles from issue to landing, so the ring has to hold > 4 slices (ncu r2b: with 4 stages the
// tensor pipe was busy 64 % of the active cycles at 37 % L2 / 31 % DRAM utilisation — waiting on
// data in flight, not on bandwidth). Evaluation launches get 5 stages by letting the stored-map
// transpose buffer share the bytes of the staged Find vectors (a tile is either fused-Find or
// stored, never both); training launches also store the Find maps for the backward pass, need
// both buffers at once, and run 4 stages.
__host__ __device__ constexpr int proj_stages(bool store_and_fuse) { return store_and_fuse ? 4 : 5; }
// Roles are split on warpgroup boundaries so that setmaxnreg can move registers from the producers
// (which need ~30) to the epilogue warps (which want > 200: a 32-column accumulator chunk in
// flight, the one being reduced, its squares, and a consumer node's two staged vectors loaded as
// ONE batch — with two epilogue warps per scheduler, load -> use -> load chains expose the
// shared-memory latency).
constexpr int kProjThreads = 384;
constexpr int kProducerRegs = 40, kEpilogueRegs = 232;   // 4*40 + 8*232 <= 2048 per 32 lanes
constexpr int kEpiThreads = 256;
constexpr int kTmemCols = 512;
// Epilogue operand staging (only when a tile spans <= 2 images, i.e. HW >= 127): for each of the
// two images up to 8 consumer nodes x (tau∘w2, tau²) x 256 columns, plus the bias. With both
// vectors staged a node costs two FMAs per column (num += m·(τw2), den += m²·τ²); the epilogue of
// a fused tile is bound by the SM's 128 fp32 lanes, so the FMA count is what matters.
constexpr int kVecImages = 2;
constexpr int kVecFloats = kVecImages * kMaxProjNodesPerPass * kBN;   // 4096 floats = 16 KB
// running (num, den) of every (row, consumer node), one copy per epilogue warpgroup (column half):
// [2 halves][8 nodes][2][128 rows]. Kept in shared memory so that the node loop is a ROLLED loop:
// unrolled over 8 nodes the epilogue was ~70 KB of straight-line code executed once per tile, and
// skipping the unused node blocks cost a chain of taken branches into cold instruction-cache lines.
constexpr int kPartFloats = 2 * kMaxProjNodesPerPass * 2 * kBM;
// stored maps leave through a per-warp 32 x 32 transpose buffer so that 8 lanes write one full
// 128-byte line (a lane owns a ROW of the accumulator: direct stores touch 32 lines per
// instruction and the L1 takes one line per cycle — measured 8.2 K cycles per stored tile)
constexpr int kStoreFloats = (kEpiThreads / 32) * 32 * 32;            // 8 warps x 4 KB
static_assert(kStoreFloats <= 2 * kVecFloats, "the store buffer aliases the staged Find vectors");
__host__ __device__ constexpr int proj_vec_bytes(bool store_and_fuse) {
  return (2 * kVecFloats + kBN + kPartFloats + (store_and_fuse ? kStoreFloats : 0)) * 4 +
         kBM * (int)sizeof(float*);
}
// dynamic smem: stages + staged vectors + barriers
__host__ __device__ constexpr int proj_smem_bytes(bool store_and_fuse) {
  return proj_stages(store_and_fuse) * kStageBytes + proj_vec_bytes(store_and_fuse) + 256;
}

struct ProjTensorMaps {
  CUtensorMap a[kMaxSeg];            // features of each segment [rows, Dk] fp32, box 32 x 128
  CUtensorMap b[NUM_PROJ_SETS];      // W^T [Mp, Kp] fp32 (K-major), box 32 x 128 (half an N-tile)
};

template <bool kStoreAndFuse>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kProjThreads, 1)
proj_umma_kernel(const __grid_constant__ ProjTensorMaps tm, const ProjParams p) {
  constexpr int kStages = proj_stages(kStoreAndFuse);
  constexpr int kRingBytes = kStages * kStageBytes;
  constexpr int kVecBytes = proj_vec_bytes(kStoreAndFuse);
  // SWIZZLE_128B tiles need 1024-byte alignment. The alignment comes from the declaration, NOT
  // from integer arithmetic on the pointer: that would demote every shared-memory access below
  // to generic LD/ST (measured: the epilogue's staged-vector reads became LD.E.128 and the
  // epilogue 2x slower).
  extern __shared__ __align__(1024) uint8_t proj_smem[];
  uint8_t* smem = proj_smem;
  if ((ptx::smem_u32(proj_smem) & 1023u) != 0) __trap();
  float* s_tw = reinterpret_cast<float*>(smem + kRingBytes);              // [2][8][256] τ∘w2
  float* s_t2 = s_tw + kVecFloats;                                         // [2][8][256] τ²
  float* s_bias = s_t2 + kVecFloats;
  float* s_part = s_bias + kBN;          // [2 halves][8 nodes][num|den][128 rows]
  // [8 warps][32 rows][32 cols], 16-byte chunks swizzled
  float* s_store = kStoreAndFuse ? s_part + kPartFloats : s_tw;
  float** s_rowdst = reinterpret_cast<float**>(s_part + kPartFloats +
                                               (kStoreAndFuse ? kStoreFloats : 0));   // [128]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kRingBytes + kVecBytes);   // leader's is used
  uint64_t* empty = full + kStages;
  uint64_t* tmem_full = empty + kStages;       // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2], leader's is used
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();        // 0 = leader of the pair
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  pdl_trigger();   // let the node kernel's CTAs start prefetching their parameters
  if (threadIdx.x == 0) N2NMN_STAMP(1, 0);

  if (warp == 0 && ptx::elect_one()) {
    for (int i = 0; i < kMaxSeg; ++i)
      if (i < p.num_seg) ptx::prefetch_tensormap(&tm.a[i]);
    for (int i = 0; i < NUM_PROJ_SETS; ++i) ptx::prefetch_tensormap(&tm.b[i]);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full[s], 1);        // the leader producer's arrive.expect_tx
      ptx::mbar_init(&empty[s], 1);       // one multicast tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], 16);   // one arrive per epilogue warp of BOTH CTAs
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc_pair<kTmemCols>(tmem_base_slot);
  ptx::tc_fence_before();
  ptx::cluster_sync_all();   // barriers of both CTAs initialised before any remote arrive / TMA
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  if (threadIdx.x == 0) N2NMN_STAMP(1, 1);

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kProducerRegs));
  if (warp == 0) {
    // ===================================================================== TMA producer
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int wi = pair; wi < p.num_work; wi += npairs) {
        const ProjWork* wk = p.work + wi;
        const int row0 = wk->row0[rank], seg = wk->seg[rank], set = wk->set;
        for (int nt = 0; nt < p.n_tiles; ++nt) {
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            ptx::mbar_wait(&empty[stage], phase ^ 1);
            // both CTAs' bytes land on the leader's barrier (the MMA issuer waits there)
            const uint32_t bar = ptx::map_to_cta(&full[stage], 0);
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full[stage], 2 * kStageBytes);
            uint8_t* dst = smem + stage * kStageBytes;
            ptx::tma_load_2d_pair(dst, &tm.a[seg], kb * kBK, row0, bar);
            ptx::tma_load_2d_pair(dst + kABytes, &tm.b[set], kb * kBK,
                                  nt * kBN + (int)rank * kBNHalf, bar);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================================================================== MMA issuer (leader)
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_tf32(2 * kBM, kBN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;   // accumulator uses so far
      for (int wi = pair; wi < p.num_work; wi += npairs) {
        for (int nt = 0; nt < p.n_tiles; ++nt, ++it) {
          const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
          ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);   // both epilogues drained this buffer
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * kBN;
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            ptx::mbar_wait(&full[stage], phase);              // TMA bytes of both CTAs landed
            if (kb < 16) N2NMN_STAMP(1, 8 + kb);
            ptx::tc_fence_after();
            const uint32_t sa = ptx::smem_u32(smem + stage * kStageBytes);
            const uint64_t da = ptx::make_smem_desc_sw128(sa);
            const uint64_t db = ptx::make_smem_desc_sw128(sa + kABytes);
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              // advance 32 bytes (= 2 x 16-byte units) along K inside the swizzle atom
              ptx::umma_tf32_pair(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            }
            ptx::umma_commit_pair(&empty[stage], 3);          // frees the slot in both CTAs
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          ptx::umma_commit_pair(&tmem_full[acc], 3);          // accumulator ready in both CTAs
        }
      }
    }
  }
  } else {
    // ===================================================================== epilogue warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpilogueRegs));
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const int trow = quarter * 32 + lane;      // row inside the 128-row tile
    const int etid = threadIdx.x - 128;        // 0..255 among the epilogue threads
    const int half = (warp - 4) >> 2;          // which 128 columns of each N-tile this warp takes
    const bool staged = p.HW >= kBM - 1;       // a tile then spans at most two images
    const uint32_t tmem_empty_leader[2] = {ptx::map_to_cta(&tmem_empty[0], 0),
                                           ptx::map_to_cta(&tmem_empty[1], 0)};
    uint32_t it = 0;
    // tauw / tau2 come from the text-projection kernel, which may still be running (PDL); the
    // TMA / MMA warps above never touch its output and start immediately.
    if (warp == 4) N2NMN_STAMP(1, 2);
    pdl_wait();
    if (warp == 4) N2NMN_STAMP(1, 3);
    for (int wi = pair; wi < p.num_work; wi += npairs) {
      const ProjWork* wkp = p.work + wi;
      const int row0 = wkp->row0[rank], pass = wkp->pass[rank];
      const int g0 = wkp->seg[rank] * p.seg_images;   // first image of this tile's segment
      struct { int set; } wk = {wkp->set};
      if (pass < 0) {
        // filler tile of an odd pair: the MMA ran on it, nothing to write — just hand the
        // accumulator back
        for (int nt = 0; nt < p.n_tiles; ++nt, ++it) {
          const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
          ptx::mbar_wait(&tmem_full[acc], acc_phase);
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(tmem_empty_leader[acc]);
        }
        continue;
      }
      const int row = row0 + trow;
      const bool row_ok = row < p.total_rows;
      const int b_first = row0 / p.HW;                 // image index inside the segment
      const int b = row_ok ? row / p.HW : b_first;
      const int pix = row - b * p.HW;
      const int img_local = b - b_first;
      // consumers of this row
      int e_beg = 0, n_nodes = 0;
      float* mdst = nullptr;
      if (row_ok) {
        if (wk.set == PS_FIND) {
          e_beg = p.img_ptr[g0 + b] + pass * kMaxProjNodesPerPass;
          n_nodes = min(p.img_ptr[g0 + b + 1] - e_beg, kMaxProjNodesPerPass);
          n_nodes = max(n_nodes, 0);
        }
        // stored map (always for the non-Find sets; for PS_FIND only in training schedules)
        const int slot = p.mslot[wk.set * p.num_images + g0 + b];
        if (slot >= 0 && pass == 0) mdst = p.mbuf + ((size_t)slot * p.HW + pix) * p.Mp;
      }
      float* acc_row = s_part + half * (kMaxProjNodesPerPass * 2 * kBM) + trow;   // + j*2*kBM (+kBM)
      const float* __restrict__ bias = p.bias[wk.set];
      const int n_img = min((p.total_rows - 1) / p.HW, (row0 + kBM - 1) / p.HW) - b_first + 1;
      // does any row of this warp's quarter leave through the stored-map path? (warp-uniform)
      const bool any_store = __any_sync(0xffffffffu, mdst != nullptr);

      for (int nt = 0; nt < p.n_tiles; ++nt, ++it) {
        // ---- stage this N-tile's epilogue operands while the MMAs are still running
        asm volatile("bar.sync 1, 256;" ::: "memory");   // previous readers of the staging are done
        if (half == 0) s_rowdst[trow] = mdst;
        for (int i = etid; i < kBN / 4; i += kEpiThreads)
          reinterpret_cast<float4*>(s_bias)[i] =
              __ldg(reinterpret_cast<const float4*>(bias + nt * kBN) + i);
        if (staged && wk.set == PS_FIND) {
          // item = (image, node, column quad): 2 x 8 x 64 float4 of tau∘w2 and of tau²
          for (int i = etid; i < kVecFloats / 4; i += kEpiThreads) {
            const int q = i & 63, j = (i >> 6) & 7, im = i >> 9;
            if (im < n_img) {
              const int eb = p.img_ptr[g0 + b_first + im] + pass * kMaxProjNodesPerPass;
              if (eb + j < p.img_ptr[g0 + b_first + im + 1]) {
                const size_t off = (size_t)p.node_text[eb + j] * p.Mp + nt * kBN;
                reinterpret_cast<float4*>(s_tw)[i] =
                    __ldg(reinterpret_cast<const float4*>(p.tauw + off) + q);
                reinterpret_cast<float4*>(s_t2)[i] =
                    __ldg(reinterpret_cast<const float4*>(p.tau2 + off) + q);
              }
            }
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp == 4) N2NMN_STAMP(1, 4);

        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kBN;
        // warp-uniform upper bound of the per-lane node counts: unused node slots are skipped by
        // a uniform branch instead of being executed predicated-off
        int n_max = n_nodes;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_max = max(n_max, __shfl_xor_sync(0xffffffffu, n_max, o));
        const int ch0 = half * (kBN / 64), ch1 = ch0 + kBN / 64;   // this warp's 4 chunks
        float4* st4 = reinterpret_cast<float4*>(s_store) + (warp - 4) * (32 * 8);
        const int cq = lane & 7;
        float* dstp[8];   // destination rows of the 8 store instructions of a chunk (per tile)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          dstp[i] = any_store ? s_rowdst[quarter * 32 + i * 4 + (lane >> 3)] : nullptr;
        ptx::mbar_wait(&tmem_full[acc], acc_phase);
        if (warp == 4) N2NMN_STAMP(1, 5);
        ptx::tc_fence_after();
        float vbuf[2][32];
        ptx::tmem_ld_32x32b_x32_nowait(taddr + ch0 * 32, vbuf[0]);
#pragma unroll 2
        for (int ch = ch0; ch < ch1; ++ch) {
          float (&v)[32] = vbuf[(ch - ch0) & 1];
          ptx::tmem_ld_wait();
          if (ch + 1 < ch1) {   // next chunk's TMEM load flies under this chunk's math
            __syncwarp();
            ptx::tmem_ld_32x32b_x32_nowait(taddr + (ch + 1) * 32, vbuf[(ch + 1 - ch0) & 1]);
          }
          const int col0 = nt * kBN + ch * 32;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bq = reinterpret_cast<const float4*>(s_bias + ch * 32)[q];
            v[4 * q + 0] += bq.x; v[4 * q + 1] += bq.y; v[4 * q + 2] += bq.z; v[4 * q + 3] += bq.w;
          }
          if (any_store) {
            // lane = row holds 32 consecutive columns; transpose through shared memory (chunk
            // index XOR row keeps both the row-wise writes and the line-wise reads conflict-free)
#pragma unroll
            for (int q = 0; q < 8; ++q)
              st4[lane * 8 + (q ^ (lane & 7))] =
                  make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            __syncwarp();
            // all eight reads first, then the stores: with two epilogue warps per scheduler a
            // load -> use -> load chain would expose the shared-memory latency eight times
            float4 tq[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {       // 4 rows x 128 bytes per instruction
              const int r = i * 4 + (lane >> 3);
              tq[i] = st4[r * 8 + (cq ^ (r & 7))];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (dstp[i] != nullptr) reinterpret_cast<float4*>(dstp[i] + col0)[cq] = tq[i];
            __syncwarp();
          }
          if (n_max > 0) {
            const bool first = (nt == 0 && ch == ch0);   // first chunk of the tile: overwrite
            if (staged) {
              // num += m·(τ∘w2) ; den += m²·τ²  — two-wide fp32 FMAs and two independent partial
              // sums per quantity; lanes whose image has fewer nodes compute into slots nobody reads
              float2 vv[16], v2[16];
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                vv[q] = make_float2(v[2 * q], v[2 * q + 1]);
                v2[q] = __fmul2_rn(vv[q], vv[q]);
              }
#pragma unroll 1
              for (int j = 0; j < n_max; ++j) {
                const int vo = (img_local * kMaxProjNodesPerPass + j) * kBN + ch * 32;
                const float4* tw = reinterpret_cast<const float4*>(s_tw + vo);
                const float4* t2 = reinterpret_cast<const float4*>(s_t2 + vo);
                // both vectors are read up front (see the note at the stored-map path)
                float4 a4[8], s4[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) a4[q] = tw[q];
#pragma unroll
                for (int q = 0; q < 8; ++q) s4[q] = t2[q];
                float* pa = acc_row + j * (2 * kBM);
                float2 na = make_float2(first ? 0.f : pa[0], 0.f), nb = make_float2(0.f, 0.f);
                float2 da = make_float2(first ? 0.f : pa[kBM], 0.f), db = nb;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  na = __ffma2_rn(vv[2 * q], make_float2(a4[q].x, a4[q].y), na);
                  nb = __ffma2_rn(vv[2 * q + 1], make_float2(a4[q].z, a4[q].w), nb);
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  da = __ffma2_rn(v2[2 * q], make_float2(s4[q].x, s4[q].y), da);
                  db = __ffma2_rn(v2[2 * q + 1], make_float2(s4[q].z, s4[q].w), db);
                }
                pa[0] = (na.x + na.y) + (nb.x + nb.y);
                pa[kBM] = (da.x + da.y) + (db.x + db.y);
              }
            } else {
              float v2[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) v2[i] = v[i] * v[i];
#pragma unroll 1
              for (int j = 0; j < n_max; ++j) {
                const int trow_txt = p.node_text[e_beg + min(j, max(n_nodes - 1, 0))];
                const float4* tw =
                    reinterpret_cast<const float4*>(p.tauw + (size_t)trow_txt * p.Mp + col0);
                const float4* t2 =
                    reinterpret_cast<const float4*>(p.tau2 + (size_t)trow_txt * p.Mp + col0);
                float* pa = acc_row + j * (2 * kBM);
                float n = first ? 0.f : pa[0], d = first ? 0.f : pa[kBM];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const float4 a = __ldg(tw + q), sq = __ldg(t2 + q);
                  n = fmaf(v[4 * q + 0], a.x, n); d = fmaf(v2[4 * q + 0], sq.x, d);
                  n = fmaf(v[4 * q + 1], a.y, n); d = fmaf(v2[4 * q + 1], sq.y, d);
                  n = fmaf(v[4 * q + 2], a.z, n); d = fmaf(v2[4 * q + 2], sq.z, d);
                  n = fmaf(v[4 * q + 3], a.w, n); d = fmaf(v2[4 * q + 3], sq.w, d);
                }
                pa[0] = n;
                pa[kBM] = d;
              }
            }
          }
          if (warp == 4) N2NMN_STAMP(1, 24 + ch);
        }
        // release the accumulator buffer to the leader's MMA warp
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(tmem_empty_leader[acc]);
      }
      if (warp == 4) N2NMN_STAMP(1, 7);
      // the two column halves of a row meet here
      if (wk.set == PS_FIND) {
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (half == 0 && row_ok) {
          const float b2 = __ldg(p.elt_b);
          const float* other = acc_row + kMaxProjNodesPerPass * 2 * kBM;
          for (int j = 0; j < n_nodes; ++j) {
            const float nn = acc_row[j * 2 * kBM] + other[j * 2 * kBM];
            const float dd = acc_row[j * 2 * kBM + kBM] + other[j * 2 * kBM + kBM];
            const int slot = p.node_out[e_beg + j];
            p.arena[(size_t)slot * p.HW + pix] = nn * rsqrtf(fmaxf(dd, kEps)) + b2;
          }
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");   // s_part is rewritten by the next tile
      }
    }
  }

  if (warp == 4) N2NMN_STAMP(1, 6);
  // teardown: the peer may still read its TMEM / the leader's MMAs may still read our operands
  __syncwarp();   // the single-lane role loops above rejoin their warps before the aligned barrier
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair<kTmemCols>(tmem_base);
  }
}

}  // namespace n2nmn
