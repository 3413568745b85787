// Tap-GEMM, CTA-pair variant (tcgen05.mma.cta_group::2, UMMA M = 256): same math and epilogue as gemm_tap.cu.
//
// Why: with cta_group::1 an SS-mode MMA reads BOTH operands from the SM's own shared memory while TMA is refilling the
// ring; for a 128 x BN tile that is (4 KB + BN*32 B) per K=16 step of operand reads plus the same again of TMA writes --
// above the 128 B/clk shared-memory port, which caps the 1-CTA kernel at ~60-70 % of the tensor peak (measured: 1.03-
// 1.25 PFLOP/s on the big 3x3 convs).  In a CTA pair each SM stages its own 128 A rows but only HALF of the B tile; the
// pair's tensor cores read the two halves from both SMs, so per-SM smem traffic for B halves.
//
// Structure (cluster of 2 CTAs on one TPC, persistent, 448 threads per CTA):
//   both CTAs : warp0 = TMA producer for its own A rows + its half of B (TMA .cta_group::2: the complete_tx lands on
//               the LEADER's full barrier), warps2..17 = epilogue for its own 128 accumulator rows (own TMEM)
//   leader    : warp1 lane0 issues tcgen05.mma.cta_group::2 for the pair; tcgen05.commit multicasts to both CTAs'
//               empty / tmem_full barriers; the leader's tmem_empty barrier collects the arrivals of BOTH epilogues
//               (the peer arrives through mapa + mbarrier.arrive.shared::cluster)
#include <cuda_runtime.h>

#include "common.cuh"
#include "gemm_common.cuh"
#include "kernels.h"

namespace vc {

template <int BN>
struct Gemm2Cfg {
  static constexpr int A_BYTES = BM * BK * 2;                  // this CTA's 128 rows, one 64-wide K sub-block
  static constexpr int B_BYTES = (BN / 2) * BK * 2;            // this CTA's half of the B tile, one K sub-block
  static constexpr int SUB_BYTES = A_BYTES + B_BYTES;
  // The single MMA-issuing thread is the bottleneck of this kernel (measured: ~200 cycles of barrier wait / fence /
  // commit per pipeline step against 320-512 cycles of tensor work), so every stage carries KSUB = 2 K sub-blocks:
  // 8 MMAs per full-barrier wait and per tcgen05.commit instead of 4.
  static constexpr int KSUB = 2;
  static constexpr int STAGE_BYTES = KSUB * SUB_BYTES;
  static constexpr int BUDGET = 227 * 1024 - 1024 - 512 - EPI_SMEM_BYTES;
  static constexpr int STAGES = BUDGET / STAGE_BYTES > 8 ? 8 : BUDGET / STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_SMEM_BYTES + 1024 + 512;
  static constexpr int NACC = (512 / BN) > 4 ? 4 : (512 / BN);   // TMEM accumulators: as many as fit the 512 columns (short-K tiles outrun two)
  static constexpr int TMEM_COLS = NACC * BN <= 32 ? 32 : NACC * BN <= 64 ? 64 : NACC * BN <= 128 ? 128 : NACC * BN <= 256 ? 256 : 512;
  static_assert(BN % 32 == 0 && (BN / 2) % 8 == 0, "B half must be whole 8-row swizzle groups");
  static_assert(B_BYTES % 1024 == 0, "B half must keep 1024-byte alignment of the next stage");
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA loads whose mbarrier completion is delivered to the pair LEADER's barrier (peer bit of the address cleared)
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma2_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit: arrive (count 1) on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1) gemm_tap2_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = Gemm2Cfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES;          // per-warp staging tiles of the TMA-store epilogue
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + EPI_SMEM_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;     // [NACC]
  uint64_t* tmem_empty_bar = tmem_full_bar + Cfg::NACC;   // [NACC]  (only the leader's copy is waited on)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + Cfg::NACC);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader
  const int cluster_id = blockIdx.x >> 1;
  const int nclusters = gridDim.x >> 1;
  const int kblocks = (p.K + BK - 1) / BK;
  const int pair_tiles = p.total_tiles;             // (m-tile pairs) x n_tiles

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_a);
    tma_prefetch_desc(&p.tmap_b);
    if (p.out_tma) tma_prefetch_desc(&p.tmap_out);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < Cfg::NACC; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 2 * EPI_WARPS);          // one arrival per epilogue warp of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // peer barriers initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ring / accumulator positions as (index, phase) pairs and multiply-high tile coordinates: see gemm_tap.cu
  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------------------
    // warp-uniform loop, one elected lane issues (keeps TMA operands in uniform registers)
    int s = 0;
    uint32_t ph = 0;
    for (int tile = cluster_id; tile < pair_tiles; tile += nclusters) {
      const int m_pair = fast_div(p.div_n_tiles, tile);
      const int n_tile = tile - m_pair * p.n_tiles;
      const TileCoord tc = tile_coord_m(p, m_pair * 2 + (int)rank);
      const int n0 = n_tile * BN + (int)rank * (BN / 2);
      for (int tap = 0; tap < p.num_taps; ++tap) {
        const int cx = tc.x0 + p.tap_dx[tap], cy = tc.y0 + p.tap_dy[tap];
        const int brow = tap * p.N + n0;
        for (int kb = 0; kb < kblocks; kb += Cfg::KSUB) {
          const int nsub = min(Cfg::KSUB, kblocks - kb);
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * nsub * Cfg::SUB_BYTES);   // bytes of both CTAs land on the leader
            for (int u = 0; u < nsub; ++u) {
              uint8_t* sa = smem + s * Cfg::STAGE_BYTES + u * Cfg::SUB_BYTES;
              uint8_t* sb = sa + Cfg::A_BYTES;
              const int k = (kb + u) * BK;
              if (k < p.K1)
                tma2_load_4d(sa, &p.tmap_a, &full_bar[s], k, cx, cy, tc.z);
              else
                tma2_load_4d(sa, &p.tmap_a2, &full_bar[s], k - p.K1, cx, cy, tc.z);
              tma2_load_2d(sb, &p.tmap_b, &full_bar[s], k, brow);
            }
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only) ------------------------------
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(2 * BM, BN);
      const uint64_t desc0 = umma_desc_sw128(smem_u32(smem));
      const uint32_t desc_hi = (uint32_t)(desc0 >> 32), desc_lo = (uint32_t)desc0;   // start-address field never carries
      int s = 0, acc = 0;
      uint32_t ph = 0, aph = 0;
      for (int tile = cluster_id; tile < pair_tiles; tile += nclusters) {
        mbar_wait(&tmem_empty_bar[acc], aph ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + acc * BN;
        uint32_t first = 1;                          // the first MMA of a tile overwrites the accumulator
        for (int tap = 0; tap < p.num_taps; ++tap) {
          for (int kb = 0; kb < kblocks; kb += Cfg::KSUB) {
            const int nsub = min(Cfg::KSUB, kblocks - kb);
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            if (elect_one()) {
              for (int u = 0; u < nsub; ++u) {
                const uint32_t la = desc_lo + (uint32_t)((s * Cfg::STAGE_BYTES + u * Cfg::SUB_BYTES) >> 4);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                  umma2_ss(tacc, ((uint64_t)desc_hi << 32) | (la + 2 * k), ((uint64_t)desc_hi << 32) | (la + (Cfg::A_BYTES >> 4) + 2 * k),
                           idesc, (first && u == 0 && k == 0) ? 0u : 1u);
              }
              umma2_commit_mc(&empty_bar[s]);
            }
            first = 0;
            __syncwarp();
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
        }
        if (elect_one()) umma2_commit_mc(&tmem_full_bar[acc]);
        __syncwarp();
        if (++acc == Cfg::NACC) { acc = 0; aph ^= 1; }
      }
    }
  } else {
    // ------------------------------ epilogue (both CTAs, own 128 rows) ------------------------------
    // the leader's tmem_empty barrier counts the epilogue warps of both CTAs
    gemm_epilogue_loop<BN, Cfg::NACC, true>(p, cluster_id, nclusters, 2, (int)rank, tmem_base, tmem_full_bar, tmem_empty_bar, epi_smem, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // nobody may exit / free TMEM while the peer can still signal it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS) : "memory");
  }
}

template <int BN>
static int launch_gemm2(const GemmParams& p, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BN>;
  static DeviceOnce configured;
  if (device_once_needed(configured)) {
    VC_CHECK_CUDA(cudaFuncSetAttribute(gemm_tap2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    device_once_mark(configured);
  }
  int nclusters = sm_count() / 2;
  if (nclusters > p.total_tiles) nclusters = p.total_tiles;
  gemm_tap2_kernel<BN><<<2 * nclusters, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(p);
  VC_CHECK_CUDA(cudaGetLastError());
  return VC_OK;
}

int launch_gemm_pair(int BN, const GemmParams& p, cudaStream_t stream) {
  switch (BN) {
    case 128: return launch_gemm2<128>(p, stream);
    case 160: return launch_gemm2<160>(p, stream);
    case 256: return launch_gemm2<256>(p, stream);
  }
  set_error("gemm_tap2: no CTA-pair kernel for BN=%d", BN);
  return VC_ERR_UNSUPPORTED;
}

}  // namespace vc
