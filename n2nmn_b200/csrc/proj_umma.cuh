// K2: conv_image contraction on the 5th-gen tensor cores (tcgen05, kind::tf32) with the module
// epilogue fused on the TMEM accumulator. See proj_common.cuh for the math and what is fused.
//
// Tiling: one work item = 128 consecutive rows of the flattened (image, pixel) axis x all Mp
// output columns, walked as Mp/256 N-tiles of 256 columns (UMMA 128x256x8, fp32 operands read as
// TF32 straight from the caller's fp32 feature grid — no conversion pass). K is streamed in
// 32-float (128-byte, one swizzle atom) slices through two TMA->smem rings, one per operand:
//     A slice 128x32 fp32 (16 KB, 3 slots) and B slice 256x32 fp32 (32 KB, 3 slots), both
//     SWIZZLE_128B K-major.
// TMEM holds two 128x256 fp32 accumulators (all 512 columns), so the epilogue of N-tile i
// overlaps the MMAs of N-tile i+1 / of the next work item. The grid is persistent: CTA i walks
// work items i, i+gridDim, ... (gridDim = min(items, SMs), or fewer on request: n2nmn_set_proj_ctas).
//
// Warp roles (384 threads): warpgroup 0 = {A-ring TMA producer, TMEM allocator + MMA issuer,
// B-ring TMA producer, idle} (one elected lane each, registers handed to the epilogue with
// setmaxnreg), warps 4..11 = epilogue; epilogue warp w reads TMEM lanes [32*(w%4), 32*(w%4)+32)
// i.e. tile rows with that offset, and the two warps that share a lane quarter split the 256
// columns of an N-tile in halves (the epilogue of a lone tile is a latency chain; two warpgroups
// halve it). Their partial sums meet in shared memory. DESIGN.md §4 has the anatomy and what was
// measured on the way.
//
// Precision: TF32 operands (10-bit mantissa), fp32 accumulate. Error budget vs the fp32/fp64
// oracle is in DESIGN.md; tests/test_gpu_parity.py holds every attention map to 1e-3 abs.
#pragma once
#include "proj_common.cuh"
#include "ptx_sm100.cuh"

namespace n2nmn {

constexpr int kBM = 128;           // rows per tile (UMMA M)
constexpr int kBN = 256;           // columns per N-tile (UMMA N)
constexpr int kBK = 32;            // fp32 elements per K slice = 128 bytes
constexpr int kUmmaK = 8;          // K per tcgen05.mma for tf32 (32 bytes)
// Operand rings: separate rings (and producer warps) for the A slices (features, 16 KB, streamed
// from DRAM exactly once) and the B slices (weights, 32 KB, L2-resident after the first touch), so
// that their depths can be chosen independently. Measured at batch 64 / 256 / 1024 (kernel alone):
// A3/B3 24.5 / 50.7 / 140 us, A5/B2 26.8 / 52.8 / 153 us: the weight ring needs the depth as much as
// the feature ring does, and 3 + 3 (144 KB) is what fits next to the epilogue staging. (Four
// 48 KB stages were ~2 us faster at batch 64 but leave no room for that staging.)
#if defined(N2NMN_EXP_STAGES_A)
constexpr int kStagesA = N2NMN_EXP_STAGES_A;
#else
constexpr int kStagesA = 3;
#endif
#if defined(N2NMN_EXP_STAGES_B)
constexpr int kStagesB = N2NMN_EXP_STAGES_B;
#else
constexpr int kStagesB = 3;
#endif
constexpr int kABytes = kBM * kBK * 4;   // 16384
constexpr int kBBytes = kBN * kBK * 4;   // 32768
constexpr int kRingBytes = kStagesA * kABytes + kStagesB * kBBytes;
// Warpgroup 0 = {A-TMA warp, MMA warp, B-TMA warp, one idle warp}, warpgroups 1-2 = epilogue. Roles are split on
// warpgroup boundaries so that setmaxnreg can move registers from the producers (which need ~30)
// to the epilogue warps (which want > 200: a 32-column accumulator chunk in flight, the one being
// reduced, its squares, and a consumer node's two staged vectors loaded as ONE batch — with two
// epilogue warps per scheduler, load -> use -> load chains expose the shared-memory latency).
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
constexpr int kVecBytes =
    (2 * kVecFloats + kBN + kPartFloats + kStoreFloats) * 4 + kBM * (int)sizeof(float*);
// dynamic smem: stages + staged vectors + barriers
constexpr int kProjSmemBytes = kRingBytes + kVecBytes + 256;

struct ProjTensorMaps {
  CUtensorMap a;                     // features [total_rows, Dk] fp32, box 32 x 128
  CUtensorMap b[NUM_PROJ_SETS];      // W^T [Mp, Kp] fp32 (K-major), box 32 x 256
};

__global__ void __launch_bounds__(kProjThreads, 1)
proj_umma_kernel(const __grid_constant__ ProjTensorMaps tm, const ProjParams p) {
  // SWIZZLE_128B tiles need 1024-byte alignment. The alignment comes from the declaration, NOT
  // from integer arithmetic on the pointer: that would demote every shared-memory access below
  // to generic LD/ST (measured: the epilogue's staged-vector reads became LD.E.128 and the
  // epilogue 2x slower).
  extern __shared__ __align__(1024) uint8_t proj_smem[];
  uint8_t* smem = proj_smem;
  if ((ptx::smem_u32(proj_smem) & 1023u) != 0) __trap();
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStagesA * kABytes;
  float* s_tw = reinterpret_cast<float*>(smem + kRingBytes);              // [2][8][256] τ∘w2
  float* s_t2 = s_tw + kVecFloats;                                         // [2][8][256] τ²
  float* s_bias = s_t2 + kVecFloats;
  float* s_part = s_bias + kBN;          // [2 halves][8 nodes][num|den][128 rows]
  float* s_store = s_part + kPartFloats; // [8 warps][32 rows][32 cols], 16-byte chunks swizzled
  float** s_rowdst = reinterpret_cast<float**>(s_store + kStoreFloats);   // [128] or nullptr
  uint64_t* full_a = reinterpret_cast<uint64_t*>(smem + kRingBytes + kVecBytes);
  uint64_t* empty_a = full_a + kStagesA;
  uint64_t* full_b = empty_a + kStagesA;
  uint64_t* empty_b = full_b + kStagesB;
  uint64_t* tmem_full = empty_b + kStagesB;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();   // let the node kernel's CTAs start prefetching their parameters
  if (threadIdx.x == 0) N2NMN_STAMP(1, 0);

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tensormap(&tm.a);
    for (int i = 0; i < NUM_PROJ_SETS; ++i) ptx::prefetch_tensormap(&tm.b[i]);
    for (int s = 0; s < kStagesA; ++s) {
      ptx::mbar_init(&full_a[s], 1);
      ptx::mbar_init(&empty_a[s], 1);
    }
    for (int s = 0; s < kStagesB; ++s) {
      ptx::mbar_init(&full_b[s], 1);
      ptx::mbar_init(&empty_b[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], 8);   // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<kTmemCols>(tmem_base_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  if (threadIdx.x == 0) N2NMN_STAMP(1, 1);

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kProducerRegs));
  if (warp == 0) {
    // ===================================================================== TMA producer, A ring
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int wi = blockIdx.x; wi < p.num_work; wi += gridDim.x) {
        const ProjWork wk = p.work[wi];
        for (int nt = 0; nt < p.n_tiles; ++nt) {
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            ptx::mbar_wait(&empty_a[stage], phase ^ 1);
            ptx::mbar_arrive_expect_tx(&full_a[stage], kABytes);
#if !defined(N2NMN_EXP_SKIP_A)      // timing experiments only (results are garbage)
            ptx::tma_load_2d(smem_a + stage * kABytes, &tm.a, kb * kBK, wk.row0, &full_a[stage]);
#else
            ptx::tma_load_2d(smem_a + stage * kABytes, &tm.a, 0, 0, &full_a[stage]);
#endif
            if (++stage == kStagesA) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================================================================== TMA producer, B ring
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int wi = blockIdx.x; wi < p.num_work; wi += gridDim.x) {
        const ProjWork wk = p.work[wi];
        for (int nt = 0; nt < p.n_tiles; ++nt) {
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            ptx::mbar_wait(&empty_b[stage], phase ^ 1);
            ptx::mbar_arrive_expect_tx(&full_b[stage], kBBytes);
#if !defined(N2NMN_EXP_SKIP_B)
            ptx::tma_load_2d(smem_b + stage * kBBytes, &tm.b[wk.set], kb * kBK, nt * kBN,
                             &full_b[stage]);
#else
            ptx::tma_load_2d(smem_b + stage * kBBytes, &tm.b[wk.set], 0, 0, &full_b[stage]);
#endif
            if (++stage == kStagesB) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_tf32(kBM, kBN);
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      uint32_t it = 0;   // accumulator uses so far
      for (int wi = blockIdx.x; wi < p.num_work; wi += gridDim.x) {
        for (int nt = 0; nt < p.n_tiles; ++nt, ++it) {
          const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
          ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);   // epilogue drained this buffer
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * kBN;
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            ptx::mbar_wait(&full_b[sb], phb);                 // TMA bytes landed
            ptx::mbar_wait(&full_a[sa], pha);
            if (kb < 16) N2NMN_STAMP(1, 8 + kb);
            ptx::tc_fence_after();
            const uint64_t da = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_a + sa * kABytes));
            const uint64_t db = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_b + sb * kBBytes));
#if !defined(N2NMN_EXP_SKIP_MMA)
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              // advance 32 bytes (= 2 x 16-byte units) along K inside the swizzle atom
              ptx::umma_tf32(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            }
#endif
            ptx::umma_commit(&empty_a[sa]);                   // frees the smem slots when done
            ptx::umma_commit(&empty_b[sb]);
            if (++sa == kStagesA) { sa = 0; pha ^= 1; }
            if (++sb == kStagesB) { sb = 0; phb ^= 1; }
          }
          ptx::umma_commit(&tmem_full[acc]);                  // accumulator ready
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
    uint32_t it = 0;
    // tauw / tau2 come from the text-projection kernel, which may still be running (PDL); the
    // TMA / MMA warps above never touch its output and start immediately.
    if (warp == 4) N2NMN_STAMP(1, 2);
    pdl_wait();
    if (warp == 4) N2NMN_STAMP(1, 3);
    for (int wi = blockIdx.x; wi < p.num_work; wi += gridDim.x) {
      const ProjWork wk = p.work[wi];
      const int row = wk.row0 + trow;
      const bool row_ok = row < p.total_rows;
      const int b_first = wk.row0 / p.HW;
      const int b = row_ok ? row / p.HW : b_first;
      const int pix = row - b * p.HW;
      const int img_local = b - b_first;
      // consumers of this row
      int e_beg = 0, n_nodes = 0;
      float* mdst = nullptr;
      if (row_ok) {
        if (wk.set == PS_FIND) {
          e_beg = p.img_ptr[b] + wk.pass * kMaxProjNodesPerPass;
          n_nodes = min(p.img_ptr[b + 1] - e_beg, kMaxProjNodesPerPass);
          n_nodes = max(n_nodes, 0);
        }
        // stored map (always for the non-Find sets; for PS_FIND only in training schedules)
        const int slot = p.mslot[wk.set * p.num_images + b];
        if (slot >= 0 && wk.pass == 0) mdst = p.mbuf + ((size_t)slot * p.HW + pix) * p.Mp;
      }
      float* acc_row = s_part + half * (kMaxProjNodesPerPass * 2 * kBM) + trow;   // + j*2*kBM (+kBM)
      const float* __restrict__ bias = p.bias[wk.set];
      const int n_img = min((p.total_rows - 1) / p.HW, (wk.row0 + kBM - 1) / p.HW) - b_first + 1;
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
              const int eb = p.img_ptr[b_first + im] + wk.pass * kMaxProjNodesPerPass;
              if (eb + j < p.img_ptr[b_first + im + 1]) {
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
#if defined(N2NMN_EXP_EPI_NOLD)
#pragma unroll
        for (int i = 0; i < 32; ++i) { vbuf[0][i] = 1.f; vbuf[1][i] = 1.f; }
#define N2NMN_TMEM_LD(addr, dst)
#define N2NMN_TMEM_WAIT()
#else
#define N2NMN_TMEM_LD(addr, dst) ptx::tmem_ld_32x32b_x32_nowait(addr, dst)
#define N2NMN_TMEM_WAIT() ptx::tmem_ld_wait()
#endif
        N2NMN_TMEM_LD(taddr + ch0 * 32, vbuf[0]);
#pragma unroll 2
        for (int ch = ch0; ch < ch1; ++ch) {
          float (&v)[32] = vbuf[(ch - ch0) & 1];
          N2NMN_TMEM_WAIT();
          if (ch + 1 < ch1) {   // next chunk's TMEM load flies under this chunk's math
            __syncwarp();
            N2NMN_TMEM_LD(taddr + (ch + 1) * 32, vbuf[(ch + 1 - ch0) & 1]);
          }
#if defined(N2NMN_EXP_EPI_NONE)
          continue;
#endif
          const int col0 = nt * kBN + ch * 32;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bq = reinterpret_cast<const float4*>(s_bias + ch * 32)[q];
            v[4 * q + 0] += bq.x; v[4 * q + 1] += bq.y; v[4 * q + 2] += bq.z; v[4 * q + 3] += bq.w;
          }
#if !defined(N2NMN_EXP_EPI_NOSTORE)
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
#endif
#if defined(N2NMN_EXP_EPI_NOMATH)
          if (false) {
#else
          if (n_max > 0) {
#endif
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
        // release the accumulator buffer to the MMA warp
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      }
      if (warp == 4) N2NMN_STAMP(1, 7);
      // the two column halves of a row meet here
      if (wk.set == PS_FIND) {
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (warp == 4) N2NMN_STAMP(1, 28);
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
        if (warp == 4) N2NMN_STAMP(1, 29);
        asm volatile("bar.sync 2, 256;" ::: "memory");   // s_part is rewritten by the next tile
        if (warp == 4) N2NMN_STAMP(1, 30);
      }
    }
  }

  if (warp == 4) N2NMN_STAMP(1, 6);
  // teardown
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace n2nmn
