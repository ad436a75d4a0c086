// K2: conv_image contraction on the 5th-gen tensor cores (tcgen05, kind::tf32) with the module
// epilogue fused on the TMEM accumulator. See proj_common.cuh for the math and what is fused.
//
// Tiling: one work item = 128 consecutive rows of the flattened (image, pixel) axis x all Mp
// output columns, walked as Mp/256 N-tiles of 256 columns (UMMA 128x256x8, fp32 operands read as
// TF32 straight from the caller's fp32 feature grid — no conversion pass). K is streamed in
// 32-float (128-byte, one swizzle atom) slices through a 4-stage TMA->smem ring:
//     stage = A tile 128x32 fp32 (16 KB) + B tile 256x32 fp32 (32 KB), both SWIZZLE_128B K-major.
// TMEM holds two 128x256 fp32 accumulators (all 512 columns), so the epilogue of N-tile i
// overlaps the MMAs of N-tile i+1 / of the next work item.
//
// Warp roles (320 threads): warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator +
// MMA issuer (one elected lane), warps 2..9 = epilogue; epilogue warp w reads TMEM lanes
// [32*(w%4), 32*(w%4)+32) i.e. tile rows with that offset, and the two warps that share a lane
// quarter split the 256 columns of an N-tile in halves (the epilogue of a lone tile is a latency
// chain; two warpgroups halve it). Their partial sums meet in shared memory.
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
#if defined(N2NMN_EXP_STAGES)
constexpr int kStages = N2NMN_EXP_STAGES;
#else
constexpr int kStages = 4;
#endif
constexpr int kABytes = kBM * kBK * 4;   // 16384
constexpr int kBBytes = kBN * kBK * 4;   // 32768
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kProjThreads = 320;       // TMA warp, MMA warp, 2 x 4 epilogue warps
constexpr int kEpiThreads = 256;
constexpr int kTmemCols = 512;
// Epilogue operand staging (only when a tile spans <= 2 images, i.e. HW >= 127): for each of the
// two images up to 8 consumer nodes x tau x 256 columns, plus conv_eltwise w2 and the bias.
// The epilogue is shared-memory-bandwidth bound (every lane = row reads every vector element), so
// only tau is staged; tau∘w2 and tau² are formed in registers.
constexpr int kVecImages = 2;
constexpr int kVecFloats = kVecImages * kMaxProjNodesPerPass * kBN;   // 4096 floats = 16 KB
// + partial (num, den) of the second epilogue warpgroup: 128 rows x 8 nodes x 2
constexpr int kPartFloats = kBM * kMaxProjNodesPerPass * 2;
constexpr int kVecBytes = (kVecFloats + 2 * kBN + kPartFloats) * 4;
// dynamic smem: stages + staged vectors + barriers
constexpr int kProjSmemBytes = kStages * kStageBytes + kVecBytes + 256;

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
  uint8_t* smem_b = smem + kStages * kABytes;
  float* s_vec = reinterpret_cast<float*>(smem + kStages * kStageBytes);
  float* s_bias = s_vec + kVecFloats;
  float* s_w2 = s_bias + kBN;
  float* s_part = s_w2 + kBN;   // [128 rows][8 nodes][2]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes + kVecBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();   // let the node kernel's CTAs start prefetching their parameters
  if (threadIdx.x == 0) N2NMN_STAMP(1, 0);

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tensormap(&tm.a);
    for (int i = 0; i < NUM_PROJ_SETS; ++i) ptx::prefetch_tensormap(&tm.b[i]);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
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

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int wi = blockIdx.x; wi < p.num_work; wi += gridDim.x) {
        const ProjWork wk = p.work[wi];
        for (int nt = 0; nt < p.n_tiles; ++nt) {
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
#if defined(N2NMN_EXP_SKIP_A)      // timing experiments only (results are garbage)
            ptx::mbar_arrive_expect_tx(&full_bar[stage], kBBytes);
            ptx::tma_load_2d(smem_b + stage * kBBytes, &tm.b[wk.set], kb * kBK, nt * kBN,
                             &full_bar[stage]);
#elif defined(N2NMN_EXP_SKIP_B)
            ptx::mbar_arrive_expect_tx(&full_bar[stage], kABytes);
            ptx::tma_load_2d(smem_a + stage * kABytes, &tm.a, kb * kBK, wk.row0,
                             &full_bar[stage]);
#else
            ptx::mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
            ptx::tma_load_2d(smem_a + stage * kABytes, &tm.a, kb * kBK, wk.row0,
                             &full_bar[stage]);
            ptx::tma_load_2d(smem_b + stage * kBBytes, &tm.b[wk.set], kb * kBK, nt * kBN,
                             &full_bar[stage]);
#endif
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_tf32(kBM, kBN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;   // accumulator uses so far
      for (int wi = blockIdx.x; wi < p.num_work; wi += gridDim.x) {
        for (int nt = 0; nt < p.n_tiles; ++nt, ++it) {
          const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
          ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);   // epilogue drained this buffer
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * kBN;
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            ptx::mbar_wait(&full_bar[stage], phase);          // TMA bytes landed
            if (kb < 16) N2NMN_STAMP(1, 8 + kb);
            ptx::tc_fence_after();
            const uint64_t da = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_a + stage * kABytes));
            const uint64_t db = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_b + stage * kBBytes));
#if !defined(N2NMN_EXP_SKIP_MMA)
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              // advance 32 bytes (= 2 x 16-byte units) along K inside the swizzle atom
              ptx::umma_tf32(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            }
#endif
            ptx::umma_commit(&empty_bar[stage]);              // frees the smem slot when done
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          ptx::umma_commit(&tmem_full[acc]);                  // accumulator ready
        }
      }
    }
  } else {
    // ===================================================================== epilogue warps
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const int trow = quarter * 32 + lane;      // row inside the 128-row tile
    const int etid = threadIdx.x - 64;         // 0..255 among the epilogue threads
    const int half = (warp - 2) >> 2;          // which 128 columns of each N-tile this warp takes
    const bool staged = p.HW >= kBM - 1;       // a tile then spans at most two images
    uint32_t it = 0;
    // tauw / tau2 come from the text-projection kernel, which may still be running (PDL); the
    // TMA / MMA warps above never touch its output and start immediately.
    if (warp == 2) N2NMN_STAMP(1, 2);
    pdl_wait();
    if (warp == 2) N2NMN_STAMP(1, 3);
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
      float num[kMaxProjNodesPerPass], den[kMaxProjNodesPerPass];
#pragma unroll
      for (int j = 0; j < kMaxProjNodesPerPass; ++j) { num[j] = 0.f; den[j] = 0.f; }
      const float* __restrict__ bias = p.bias[wk.set];
      const int n_img = min((p.total_rows - 1) / p.HW, (wk.row0 + kBM - 1) / p.HW) - b_first + 1;

      for (int nt = 0; nt < p.n_tiles; ++nt, ++it) {
        // ---- stage this N-tile's epilogue operands while the MMAs are still running
        asm volatile("bar.sync 1, 256;" ::: "memory");   // previous readers of s_vec are done
        for (int i = etid; i < kBN / 4; i += kEpiThreads)
          reinterpret_cast<float4*>(s_bias)[i] =
              __ldg(reinterpret_cast<const float4*>(bias + nt * kBN) + i);
        if (staged && wk.set == PS_FIND) {
          for (int i = etid; i < kBN / 4; i += kEpiThreads)
            reinterpret_cast<float4*>(s_w2)[i] =
                (nt * kBN + i * 4 < p.M)   // conv_eltwise weights are [M], not padded
                    ? make_float4(p.elt_w[min(nt * kBN + i * 4 + 0, p.M - 1)],
                                  nt * kBN + i * 4 + 1 < p.M ? p.elt_w[nt * kBN + i * 4 + 1] : 0.f,
                                  nt * kBN + i * 4 + 2 < p.M ? p.elt_w[nt * kBN + i * 4 + 2] : 0.f,
                                  nt * kBN + i * 4 + 3 < p.M ? p.elt_w[nt * kBN + i * 4 + 3] : 0.f)
                    : make_float4(0.f, 0.f, 0.f, 0.f);
          // item = (image, node, column quad): 2 x 8 x 64 float4 of tau
          for (int i = etid; i < kVecFloats / 4; i += kEpiThreads) {
            const int q = i & 63, j = (i >> 6) & 7, im = i >> 9;
            if (im < n_img) {
              const int eb = p.img_ptr[b_first + im] + wk.pass * kMaxProjNodesPerPass;
              if (eb + j < p.img_ptr[b_first + im + 1]) {
                const float* src = p.tau + (size_t)p.node_text[eb + j] * p.Mp + nt * kBN;
                reinterpret_cast<float4*>(s_vec)[i] = __ldg(reinterpret_cast<const float4*>(src) + q);
              }
            }
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp == 2) N2NMN_STAMP(1, 4);

        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kBN;
        // warp-uniform upper bound of the per-lane node counts: unused node slots are skipped by
        // a uniform branch instead of being executed predicated-off
        int n_max = n_nodes;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_max = max(n_max, __shfl_xor_sync(0xffffffffu, n_max, o));
        const int ch0 = half * (kBN / 64), ch1 = ch0 + kBN / 64;   // this warp's 4 chunks
        ptx::mbar_wait(&tmem_full[acc], acc_phase);
        if (warp == 2) N2NMN_STAMP(1, 5);
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
          if (mdst != nullptr) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              reinterpret_cast<float4*>(mdst + col0)[q] =
                  make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
#endif
#if defined(N2NMN_EXP_EPI_NOMATH)
          if (false) {
#else
          if (n_max > 0) {
#endif
            if (staged) {
              // num += Σ (m·w2)·tau ; den += Σ (m·tau)²  — one vector (tau) read per node.
              // Two-wide fp32 FMAs (FFMA2) and two independent partial sums per quantity: with one
              // epilogue warp per scheduler the loop is otherwise bound by the FMA dependency chain.
              float2 vw[16], vv[16];
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 w4 = reinterpret_cast<const float4*>(s_w2 + ch * 32)[q];
                vv[2 * q] = make_float2(v[4 * q], v[4 * q + 1]);
                vv[2 * q + 1] = make_float2(v[4 * q + 2], v[4 * q + 3]);
                vw[2 * q] = __fmul2_rn(vv[2 * q], make_float2(w4.x, w4.y));
                vw[2 * q + 1] = __fmul2_rn(vv[2 * q + 1], make_float2(w4.z, w4.w));
              }
#pragma unroll
              for (int j = 0; j < kMaxProjNodesPerPass; ++j) {
                if (j < n_max) {            // warp-uniform
                  const float4* tv = reinterpret_cast<const float4*>(
                      s_vec + (img_local * kMaxProjNodesPerPass + j) * kBN + ch * 32);
                  float2 na = make_float2(0.f, 0.f), nb = na, da = na, db = na;
#pragma unroll
                  for (int q = 0; q < 8; ++q) {
                    const float4 t4 = tv[q];
                    const float2 ta = make_float2(t4.x, t4.y), tb2 = make_float2(t4.z, t4.w);
                    na = __ffma2_rn(vw[2 * q], ta, na);
                    nb = __ffma2_rn(vw[2 * q + 1], tb2, nb);
                    const float2 ea = __fmul2_rn(vv[2 * q], ta), eb = __fmul2_rn(vv[2 * q + 1], tb2);
                    da = __ffma2_rn(ea, ea, da);
                    db = __ffma2_rn(eb, eb, db);
                  }
                  if (j < n_nodes) {        // lanes with fewer nodes discard
                    num[j] += (na.x + na.y) + (nb.x + nb.y);
                    den[j] += (da.x + da.y) + (db.x + db.y);
                  }
                }
              }
            } else {
              float v2[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) v2[i] = v[i] * v[i];
#pragma unroll
              for (int j = 0; j < kMaxProjNodesPerPass; ++j) {
                if (j < n_max) {            // warp-uniform
                  const int trow_txt = p.node_text[e_beg + min(j, max(n_nodes - 1, 0))];
                  const float4* tw =
                      reinterpret_cast<const float4*>(p.tauw + (size_t)trow_txt * p.Mp + col0);
                  const float4* t2 =
                      reinterpret_cast<const float4*>(p.tau2 + (size_t)trow_txt * p.Mp + col0);
                  float n = num[j], d = den[j];
#pragma unroll
                  for (int q = 0; q < 8; ++q) {
                    const float4 a = tw[q], sq = t2[q];
                    n = fmaf(v[4 * q + 0], a.x, n); d = fmaf(v2[4 * q + 0], sq.x, d);
                    n = fmaf(v[4 * q + 1], a.y, n); d = fmaf(v2[4 * q + 1], sq.y, d);
                    n = fmaf(v[4 * q + 2], a.z, n); d = fmaf(v2[4 * q + 2], sq.z, d);
                    n = fmaf(v[4 * q + 3], a.w, n); d = fmaf(v2[4 * q + 3], sq.w, d);
                  }
                  if (j < n_nodes) { num[j] = n; den[j] = d; }
                }
              }
            }
          }
          if (warp == 2) N2NMN_STAMP(1, 24 + ch);
        }
        // release the accumulator buffer to the MMA warp
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      }
      if (warp == 2) N2NMN_STAMP(1, 7);
      // the two column halves of a row meet here: warpgroup 1 hands its partial sums over
      if (wk.set == PS_FIND) {
        if (half == 1) {
#pragma unroll
          for (int j = 0; j < kMaxProjNodesPerPass; ++j) {
            s_part[(trow * kMaxProjNodesPerPass + j) * 2] = num[j];
            s_part[(trow * kMaxProjNodesPerPass + j) * 2 + 1] = den[j];
          }
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (warp == 2) N2NMN_STAMP(1, 28);
        if (half == 0 && row_ok) {
          const float b2 = __ldg(p.elt_b);
#pragma unroll
          for (int j = 0; j < kMaxProjNodesPerPass; ++j) {
            if (j < n_nodes) {
              const float nn = num[j] + s_part[(trow * kMaxProjNodesPerPass + j) * 2];
              const float dd = den[j] + s_part[(trow * kMaxProjNodesPerPass + j) * 2 + 1];
              const int slot = p.node_out[e_beg + j];
              p.arena[(size_t)slot * p.HW + pix] = nn * rsqrtf(fmaxf(dd, kEps)) + b2;
            }
          }
        }
        if (warp == 2) N2NMN_STAMP(1, 29);
        asm volatile("bar.sync 2, 256;" ::: "memory");   // s_part may be rewritten by the next tile
        if (warp == 2) N2NMN_STAMP(1, 30);
      }
    }
  }

  if (warp == 2) N2NMN_STAMP(1, 6);
  // teardown
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace n2nmn
