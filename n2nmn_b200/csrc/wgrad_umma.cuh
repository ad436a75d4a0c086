// Weight gradient of the feature-side layers on the 5th-gen tensor cores:
//   dW_set[k, c] += Σ_entries Σ_p X_b[p, k] · B_entry[p, c]        (backward.cuh, FeatGradSrc)
// as tcgen05.mma kind::tf32 with BOTH operands MN-major: the contraction index p (pixels) is the
// slow index of X [p][k] and of the B maps [p][c], which is exactly how the forward pass and the
// reverse walk leave them in memory — no transposed copy of either. TMA boxes of 32 pixels x 32
// elements (128 bytes, SWIZZLE_128B_ATOM_32B) are the canonical MN-major blocks of 4-byte
// operands (ptx_sm100.cuh; the layout was established with tools/exp/umma_mn_test.cu); pixels
// beyond H·W are zero-filled by the tensor maps (3-D: element, pixel, image / entry).
//   CTA = one 128-feature slab of k x all 256 channels x a chunk of the entries (sorted by weight
//   set by the host): M = 128, N = 256, K = 8 per instruction, 4 instructions per 32-pixel stage,
//   4-stage ring of 48 KB; the accumulator (128 lanes x 256 fp32 columns of TMEM) is flushed with
//   float2 reductions when the weight set changes and at the end.
//   warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = flush.
// Operands are the fp32 bits read as TF32 (as in the forward contraction); fp32 accumulate.
#pragma once
#include <cstdio>
#include "backward.cuh"
#include "ptx_sm100.cuh"

namespace n2nmn {

constexpr int kWgM = 128, kWgN = 256, kWgP = 32, kWgStages = 4, kWgThreads = 192;
constexpr int kWgBlockBytes = kWgP * 128;                 // one [32 pixels][32 elements] box
constexpr int kWgABytes = (kWgM / 32) * kWgBlockBytes;    // 16 KB
constexpr int kWgBBytes = (kWgN / 32) * kWgBlockBytes;    // 32 KB
constexpr int kWgStageBytes = kWgABytes + kWgBBytes;
constexpr int kWgFlushFloats = 4 * 32 * 33;   // per flush warp: a 32 x 32 transpose tile (pitch 33)
constexpr size_t kWgSmemBytes = (size_t)kWgStages * kWgStageBytes + 256 + kWgFlushFloats * sizeof(float);

// 4-D views with the 32-element blocks as their own dimension, so that ONE box per operand and
// stage lands as [block][pixel][32 elements] = the canonical layout (12 boxes of 4 KB per stage
// from 3-D maps kept the single producer thread busier than the tensor core).
struct WgradMaps {
  CUtensorMap x;    // features (32, p: HW, k/32: Dk/32, image: N)       box (32, 32, 4, 1)
  CUtensorMap b;    // B maps   (32, p: HW, c/32: Mp/32, entry: E_cap)   box (32, 32, 8, 1)
};
struct WgradParams {
  const BwdEntry* entries;     // [num_entries] {set, image}
  const int32_t* order;        // entry indices sorted by weight set
  int num_entries, per_cta, HW, Dk, M;
  float* gflat;
  GradOffsets go;
};

// (A 4-CTA cluster variant that multicast the B stage to the four feature slabs — a quarter of the
// B-map traffic — was built and measured: no faster, 36.8 vs 32.7 us; the kernel was bound by the
// reductions of its flush, see below, not by operand traffic. Removed.)
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_umma_kernel(const __grid_constant__ WgradMaps tm, const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t wg_smem[];
  if ((ptx::smem_u32(wg_smem) & 1023u) != 0) __trap();
  uint64_t* full = reinterpret_cast<uint64_t*>(wg_smem + kWgStages * kWgStageBytes);
  uint64_t* empty = full + kWgStages;
  uint64_t* tmem_full = empty + kWgStages;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
  float* s_flush = reinterpret_cast<float*>(wg_smem + kWgStages * kWgStageBytes + 256);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * kWgM;
  const int e0 = blockIdx.y * p.per_cta, e1 = min(p.num_entries, e0 + p.per_cta);
  const int stages_per_entry = (p.HW + kWgP - 1) / kWgP;

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tensormap(&tm.x);
    ptx::prefetch_tensormap(&tm.b);
    for (int s = 0; s < kWgStages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    ptx::mbar_init(tmem_full, 1);
    ptx::mbar_init(tmem_empty, 4);    // one arrive per flush warp
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<kWgN>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    // ================================================================= TMA producer
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = e0; i < e1; ++i) {
        const int e = p.order[i], img = p.entries[e].b;
        for (int ps = 0; ps < stages_per_entry; ++ps) {
          ptx::mbar_wait_bounded(&empty[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full[stage], kWgStageBytes);
          uint8_t* sa = wg_smem + stage * kWgStageBytes;
          uint8_t* sb = sa + kWgABytes;
          // one box per operand: (32 elements, 32 pixels, 4 or 8 element blocks, 1 image / entry)
          ptx::tma_load_4d(sa, &tm.x, 0, ps * kWgP, k0 / 32, img, &full[stage]);
          ptx::tma_load_4d(sb, &tm.b, 0, ps * kWgP, 0, e, &full[stage]);
          if (++stage == kWgStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (ptx::elect_one()) {
#if defined(N2NMN_EXP_WG_IDESC)
      constexpr uint32_t idesc = ptx::make_idesc_tf32(kWgM, kWgN) | (N2NMN_EXP_WG_IDESC);
#else
      constexpr uint32_t idesc = ptx::make_idesc_tf32_mn(kWgM, kWgN);
#endif
      int stage = 0, cur_set = -1;
      uint32_t phase = 0, flushes = 0;
      for (int i = e0; i < e1; ++i) {
        const int set = p.entries[p.order[i]].set;
        bool fresh = false;
        if (set != cur_set) {
          if (cur_set >= 0) {
            ptx::umma_commit(tmem_full);                        // accumulator of the old set ready
            ptx::mbar_wait_bounded(tmem_empty, flushes & 1);    // ... and drained
            ++flushes;
            ptx::tc_fence_after();
          }
          cur_set = set;
          fresh = true;
        }
        for (int ps = 0; ps < stages_per_entry; ++ps) {
          ptx::mbar_wait_bounded(&full[stage], phase);
          ptx::tc_fence_after();
#if defined(N2NMN_EXP_WGDBG)
          if (blockIdx.x == 0 && blockIdx.y == 0 && i == e0 && ps == 0) {
            const float* fa = reinterpret_cast<const float*>(wg_smem + stage * kWgStageBytes);
            const float* fb = fa + kWgABytes / 4;
            printf("wgdbg e=%d img=%d set=%d A[0..3]=%g %g %g %g A[row1]=%g %g B[0..3]=%g %g %g %g B[blk1]=%g\n",
                   p.order[i], p.entries[p.order[i]].b, set, fa[0], fa[1], fa[2], fa[3], fa[32], fa[33],
                   fb[0], fb[1], fb[2], fb[3], fb[1024]);
          }
#endif
          const uint32_t a_addr = ptx::smem_u32(wg_smem + stage * kWgStageBytes);
          const uint32_t b_addr = a_addr + kWgABytes;
#pragma unroll
          for (int j = 0; j < kWgP / 8; ++j) {   // 8 pixels = two 512-byte atoms per block
            const uint64_t da = ptx::make_smem_desc_sw128_mn(a_addr + j * 1024, kWgBlockBytes, 512);
            const uint64_t db = ptx::make_smem_desc_sw128_mn(b_addr + j * 1024, kWgBlockBytes, 512);
            ptx::umma_tf32(tmem_d, da, db, idesc, !(fresh && ps == 0 && j == 0));
          }
          ptx::umma_commit(&empty[stage]);
          if (++stage == kWgStages) { stage = 0; phase ^= 1; }
        }
      }
      if (cur_set >= 0) ptx::umma_commit(tmem_full);
    }
  } else {
    // ================================================================= flush warps (TMEM -> gflat)
    const int quarter = warp & 3;                 // the TMEM lane quarter this warp may read
    float* tile = s_flush + quarter * 32 * 33;
    int cur_set = -1;
    uint32_t flushes = 0;
    // TMEM hands a lane one ROW (feature) and 32 columns; reductions issued that way touch 32
    // different lines per instruction and the L2's reduction rate set the kernel's time (23 of 33
    // us: the time did not move with the number of entry chunks). Each 32 x 32 chunk therefore goes
    // through a transpose tile so that half a warp covers 128 contiguous bytes of one row.
    auto flush = [&](int set) {
      ptx::mbar_wait_bounded(tmem_full, flushes & 1);
      ptx::tc_fence_after();
      float* W = p.gflat + p.go.proj_w[set];
      const bool pair_ok = (p.M & 1) == 0 && (reinterpret_cast<uintptr_t>(W) & 7) == 0;
      const int sub = lane >> 4, c2 = 2 * (lane & 15);
      for (int cb = 0; cb < kWgN; cb += 32) {
        float v[32];
        ptx::tmem_ld_32x32b_x32(tmem_d + (static_cast<uint32_t>(quarter * 32) << 16) + cb, v);
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 32; ++c) tile[lane * 33 + c] = v[c];
        __syncwarp();
        const int col = cb + c2;
#pragma unroll 4
        for (int rr = 0; rr < 32; rr += 2) {
          const int row = rr + sub, k = k0 + quarter * 32 + row;
          const float a = tile[row * 33 + c2], b = tile[row * 33 + c2 + 1];
          if (k >= p.Dk) continue;
          float* dst = W + (size_t)k * p.M + col;
          if (pair_ok && col + 1 < p.M) {
            if (a != 0.f || b != 0.f) atomicAdd(reinterpret_cast<float2*>(dst), make_float2(a, b));
          } else {
            if (col < p.M && a != 0.f) atomicAdd(dst, a);
            if (col + 1 < p.M && b != 0.f) atomicAdd(dst + 1, b);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tmem_empty);
      ++flushes;
    };
    for (int i = e0; i < e1; ++i) {
      const int set = p.entries[p.order[i]].set;
      if (set != cur_set) {
        if (cur_set >= 0) flush(cur_set);
        cur_set = set;
      }
    }
    if (cur_set >= 0) flush(cur_set);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kWgN>(tmem_d);
  }
}

// Bias gradient of the same layers: db_set[c] += Σ_p B_entry[p, c]. One CTA per B map; 4 row
// groups x 256 columns so that every thread has ~38 independent loads (one thread per column over
// all 150 rows was a 26 us latency chain).
__global__ void __launch_bounds__(1024)
bmap_colsum_kernel(const float* __restrict__ dmap, const BwdEntry* __restrict__ entries, int HW,
                   int M, int Mp, float* __restrict__ gflat, GradOffsets go) {
  __shared__ float part[4][256];
  const int e = blockIdx.x, c = threadIdx.x & 255, rg = threadIdx.x >> 8;
  const float* B = dmap + (size_t)e * HW * Mp;
  float* dst = gflat + go.proj_b[entries[e].set];
  for (int c0 = 0; c0 < M; c0 += 256) {
    const int col = c0 + c;
    float s = 0.f;
    if (col < M) {
#pragma unroll 8
      for (int p = rg; p < HW; p += 4) s += B[(size_t)p * Mp + col];
    }
    part[rg][c] = s;
    __syncthreads();
    if (rg == 0 && col < M) {
      const float t = (part[0][c] + part[1][c]) + (part[2][c] + part[3][c]);
      if (t != 0.f) atomicAdd(dst + col, t);
    }
    __syncthreads();
  }
}

}  // namespace n2nmn
