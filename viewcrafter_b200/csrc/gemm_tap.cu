// Tap-GEMM on tcgen05: D[M,N] = sum_taps A_tap[M,K] * W_tap[N,K]^T  (fp16 in, fp32 accumulate in TMEM).
//
// One kernel family covers every tensor-core op on the U-Net / VAE path:
//   * nn.Linear / 1x1 conv            : 1 tap, A = [M,K] row-major
//   * 3x3 conv, stride 1, pad 1 (NHWC): 9 taps, A tile = TMA box (64ch, bw, bh, 1 frame) shifted by (dx-1, dy-1);
//                                       the zero padding is TMA out-of-bounds fill
//   * Conv3d (3,1,1), pad (1,0,0)     : 3 taps, A rows shifted by +-H*W rows of the [T*H*W, C] matrix (OOB rows = 0)
// A may come from two tensors split along K (channel concat of skip connections without materialising it).
// Epilogue: + bias[z/bias_z_div][n], GEGLU (value*gelu(gate)), + residual, fp16 / fp32 store (gemm_common.cuh).
//
// This file: the host entry point and the ONE-CTA-per-tile persistent kernel (64 + 32 * EPI_WARPS = 448 threads):
//   warp 0       TMA producer: walks this CTA's tiles and their (tap, k-block) iterations through a smem ring without
//                draining between tiles, so the loads of tile i+1 are in flight while tile i is still being multiplied
//   warp 1       TMEM allocator + single-thread tcgen05.mma issuer; TWO accumulators in TMEM (double buffer), so the
//                main loop of tile i+1 overlaps the epilogue of tile i
//   warps 2..    epilogue (gemm_epilogue_loop, gemm_common.cuh)
// Tiles are ordered n-fastest so CTAs that run concurrently share A tiles in L2.  Large problems are routed to the
// CTA-pair kernel of gemm_tap2.cu (UMMA M=256), which halves the per-SM shared-memory traffic for B.
#include <cstdlib>

#include "gemm_common.cuh"
#include "kernels.h"

namespace vc {

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BUDGET = 227 * 1024 - 1024 /*align slack*/ - 512 /*barriers*/ - EPI_SMEM_BYTES /*store staging*/;
  static constexpr int STAGES = BUDGET / STAGE_BYTES > 8 ? 8 : BUDGET / STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_SMEM_BYTES + 1024 + 512;
  // as many accumulators as fit the 512 TMEM columns (2..4): short-K tiles finish their main loop faster than the
  // accumulator hand-off (commit -> epilogue wake-up -> drain -> arrive) can turn around, so two buffers are not enough
  static constexpr int NACC = (512 / BN) > 4 ? 4 : (512 / BN);
  static constexpr int TMEM_COLS = NACC * BN <= 32 ? 32 : NACC * BN <= 64 ? 64 : NACC * BN <= 128 ? 128 : NACC * BN <= 256 ? 256 : 512;
  static_assert(NACC >= 2 && NACC * BN <= 512, "at least two accumulators must fit TMEM");
  static_assert(STAGES >= 4, "pipeline too shallow");
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tap_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES;          // per-warp staging tiles of the TMA-store epilogue
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + EPI_SMEM_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;     // [NACC]
  uint64_t* tmem_empty_bar = tmem_full_bar + Cfg::NACC;   // [NACC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + Cfg::NACC);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kblocks = (p.K + BK - 1) / BK;
  const int iters = p.num_taps * kblocks;

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
      mbar_init(&tmem_empty_bar[a], EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Ring / accumulator positions are carried as (index, phase bit) pairs and tile coordinates come from multiply-high
  // divisions: the role loops below are single-warp instruction streams whose length bounds the tile rate of short-K
  // problems (measured: with 64-bit `it % STAGES` arithmetic and generic divisions the empty skeleton -- no TMA, no MMA,
  // no epilogue -- already cost 1.3 us per tile).
  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    // The whole warp runs the loop with warp-uniform control flow and ONE elected lane issues: the compiler can then
    // keep barrier / descriptor operands in uniform registers (a divergent `if (lane == 0)` loop costs an ELECT/BRA.ANY
    // serialisation loop around every UTMALDG -- measured).
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int m_tile = fast_div(p.div_n_tiles, tile);
      const TileCoord tc = tile_coord_m(p, m_tile);
      const int n0 = (tile - m_tile * p.n_tiles) * BN;
      for (int tap = 0; tap < p.num_taps; ++tap) {
        const int cx = tc.x0 + p.tap_dx[tap], cy = tc.y0 + p.tap_dy[tap];
        const int brow = tap * p.N + n0;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);        // a fresh barrier passes the wait on the "previous" phase
          if (elect_one()) {
            uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
            uint8_t* sb = sa + Cfg::A_BYTES;
            if (VC_GEMM_DBG(p, 2)) {
              mbar_arrive(&full_bar[s]);
            } else {
              mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
              const int k = kb * BK;
              if (k < p.K1)
                tma_load_4d(sa, &p.tmap_a, &full_bar[s], k, cx, cy, tc.z);
              else
                tma_load_4d(sa, &p.tmap_a2, &full_bar[s], k - p.K1, cx, cy, tc.z);
              tma_load_2d(sb, &p.tmap_b, &full_bar[s], k, brow);
            }
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    // warp-uniform loop, one elected lane issues (always the same lane: tcgen05.commit tracks the issuing thread's MMAs)
    constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
    // smem descriptors: only the 14-bit start-address field (bytes >> 4) changes between stages / k-slices and the ring
    // lies below 256 KB, so the field never carries: descriptors are formed by integer adds on the low word
    const uint64_t desc0 = umma_desc_sw128(smem_u32(smem));
    const uint32_t desc_hi = (uint32_t)(desc0 >> 32), desc_lo = (uint32_t)desc0;
    int s = 0, acc = 0;
    uint32_t ph = 0, aph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty_bar[acc], aph ^ 1);    // the epilogue must have drained this accumulator
      tc_fence_after();
      const uint32_t tacc = tmem_base + acc * BN;
      for (int i = 0; i < iters; ++i) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          if (VC_GEMM_DBG(p, 1)) {
            mbar_arrive(&empty_bar[s]);
          } else {
            const uint32_t la = desc_lo + (uint32_t)(s * (Cfg::STAGE_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_ss(tacc, ((uint64_t)desc_hi << 32) | (la + 2 * k), ((uint64_t)desc_hi << 32) | (la + (Cfg::A_BYTES >> 4) + 2 * k), idesc,
                      (i > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty_bar[s]);   // frees the smem stage once these MMAs have read it
          }
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      if (elect_one()) {
        if (VC_GEMM_DBG(p, 1)) mbar_arrive(&tmem_full_bar[acc]);
        else umma_commit(&tmem_full_bar[acc]);   // accumulator complete
      }
      __syncwarp();
      if (++acc == Cfg::NACC) { acc = 0; aph ^= 1; }
    }
  } else {
    // ------------------------------ epilogue ------------------------------
    gemm_epilogue_loop<BN, Cfg::NACC, false>(p, blockIdx.x, gridDim.x, 1, 0, tmem_base, tmem_full_bar, tmem_empty_bar, epi_smem, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN>
static int launch_gemm(const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static DeviceOnce configured;
  if (device_once_needed(configured)) {
    VC_CHECK_CUDA(cudaFuncSetAttribute(gemm_tap_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    device_once_mark(configured);
  }
  const int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
  gemm_tap_kernel<BN><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(p);
  VC_CHECK_CUDA(cudaGetLastError());
  return VC_OK;
}

static int pick_bn(int N, int geglu) {
  if (geglu) {
    static int g = -1;                       // tuning switch VC_GEGLU_BN=128|256
    if (g < 0) { const char* e = getenv("VC_GEGLU_BN"); g = e ? atoi(e) : 256; }
    return (g == 256 && N % 256 == 0) ? 256 : 128;
  }
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (N % 256 == 0) return 256;
  if (N % 160 == 0) return 160;
  if (N % 128 == 0) return 128;
  if (N % 96 == 0 && N <= 192) return 96;
  if (N % 64 == 0 && N < 256) return 64;
  return 128;
}

int pick_bn_public(int N, int geglu) { return pick_bn(N, geglu); }

// tuning switch: VC_GEMM_PAIR=0 forces the one-CTA kernel everywhere, =1 (default) uses CTA pairs for large problems
static int pair_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("VC_GEMM_PAIR");
    mode = (e && e[0] == '0') ? 0 : 1;
  }
  return mode;
}

int gemm_tap(const GemmDesc& d, cudaStream_t stream) {
  VC_REQUIRE(d.a && d.w && (d.out || d.out_f32), "gemm_tap: null pointer");
  VC_REQUIRE(d.num_taps >= 1 && d.num_taps <= MAX_TAPS, "gemm_tap: num_taps=%d out of range", d.num_taps);
  VC_REQUIRE(d.bx * d.by == BM && d.bx >= 1, "gemm_tap: box %dx%d must cover 128 rows", d.bx, d.by);
  VC_REQUIRE(d.by == 1 || d.bx == d.X, "gemm_tap: multi-row boxes need bx == X (X=%d bx=%d)", d.X, d.bx);
  VC_REQUIRE(d.K % 8 == 0 && d.lda % 8 == 0, "gemm_tap: K and lda must be multiples of 8 (TMA 16-byte strides)");
  VC_REQUIRE(d.K1 == d.K || (d.a2 && d.K1 % BK == 0 && d.K1 < d.K), "gemm_tap: bad K split K1=%d K=%d", d.K1, d.K);
  VC_REQUIRE(!d.geglu || (d.N % 128 == 0 && !d.res && !d.out_f32), "gemm_tap: GEGLU needs N %% 128 == 0");
  if ((reinterpret_cast<uintptr_t>(d.a) & 15) || (reinterpret_cast<uintptr_t>(d.w) & 15)) {
    set_error("gemm_tap: operands must be 16-byte aligned");
    return VC_ERR_ARG;
  }
  const void* optr = d.out_f32 ? (const void*)d.out_f32 : (const void*)d.out;
  const int esz = d.out_f32 ? 4 : 2;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  int BN = pick_bn(d.N, d.geglu);
  int tuned = 0;                              // 0: default pair rule; 1: tuned -> CTA pairs; 2: tuned -> single CTAs
  {
    // Wave quantisation: the persistent grid runs ceil(tiles / slots) rounds.  A frame shard of a multi-GPU run (7 frames at
    // 18x32: 32 row tiles x 1280 columns) gives 80 pair tiles on 74 SM pairs -- two rounds, the second 8 % full.  When the default
    // tiling fills its rounds to less than 80 %, compare (pair | single CTA) x (BN 256 | 160 | 128) by
    // rounds x tile area / relative tile efficiency and take the cheapest.  GEGLU keeps the tile its weights were interleaved for.
    static int tune = -1;                     // tuning switch VC_GEMM_WAVE_TUNE=0 keeps the fixed choice
    if (tune < 0) { const char* e = getenv("VC_GEMM_WAVE_TUNE"); tune = (e && e[0] == '0') ? 0 : 1; }
    const long long mt = (long long)((d.X + d.bx - 1) / d.bx) * ((d.Y + d.by - 1) / d.by) * d.Z;
    const int k_it = d.num_taps * ((d.K + BK - 1) / BK);
    const int sms = sm_count();
    auto cost = [&](int bn, bool pair, double* fill) -> double {
      const long long nt = (d.N + bn - 1) / bn;
      const long long tiles = (pair ? (mt + 1) / 2 : mt) * nt;
      const long long slots = pair ? sms / 2 : sms;
      const long long rounds = (tiles + slots - 1) / slots;
      if (fill) *fill = (double)tiles / (double)(rounds * slots);
      const double eff = (pair ? 1.08 : 1.0) * (bn == 256 ? 1.0 : bn == 160 ? 0.97 : 0.94);
      return (double)rounds * (pair ? 2.0 : 1.0) * bn / eff;
    };
    auto pair_ok = [&](int bn) { return pair_mode() && d.N % bn == 0 && k_it >= 5 && (mt / 2) * ((d.N + bn - 1) / bn) >= 1; };
    if (tune && !d.geglu && d.N >= 128 && (BN == 256 || BN == 160 || BN == 128)) {
      double fill = 1.0;
      const bool def_pair = pair_ok(BN) && (mt / 2) * ((d.N + BN - 1) / BN) >= sms;
      double best = cost(BN, def_pair, &fill);
      if (fill < 0.8) {
        int best_bn = BN; bool best_pair = def_pair;
        const int cands[3] = {256, 160, 128};
        for (int ci = 0; ci < 3; ++ci) {
          const int bn = cands[ci];
          if (d.N % bn != 0) continue;
          for (int pr = 0; pr < 2; ++pr) {
            if (pr && !pair_ok(bn)) continue;
            const double c = cost(bn, pr != 0, nullptr);
            if (c < best * 0.97) { best = c; best_bn = bn; best_pair = pr != 0; }
          }
        }
        BN = best_bn;
        tuned = best_pair ? 1 : 2;
      }
    }
  }
  p.bx = d.bx; p.by = d.by; p.X = d.X; p.Y = d.Y; p.Z = d.Z;
  p.tiles_x = (d.X + d.bx - 1) / d.bx;
  p.tiles_y = (d.Y + d.by - 1) / d.by;
  p.n_tiles = (d.N + BN - 1) / BN;
  p.div_tiles_x = make_fastdiv(p.tiles_x);
  p.div_tiles_y = make_fastdiv(p.tiles_y);
  p.div_n_tiles = make_fastdiv(p.n_tiles);
  p.bx_shift = 0;
  while ((1 << p.bx_shift) < d.bx) ++p.bx_shift;
  VC_REQUIRE((1 << p.bx_shift) == d.bx, "gemm_tap: bx=%d must be a power of two", d.bx);
  const long long m_tiles = (long long)p.tiles_x * p.tiles_y * p.Z;
  // CTA pairs pay off once every SM pair has several 256-row tiles; N must be covered by whole BN tiles so that each
  // CTA's half of the B tile (BN/2 rows) never straddles a tap boundary
  // Measured on B200 (profiles/README.md): pairs win 10-15 % on long reductions; since the role loops were slimmed down they
  // also win 4-13 % on the short-K (K = 320 / 512, 5-8 k-blocks) level-0 linears (they lost there before).
  const int k_iters = d.num_taps * ((d.K + BK - 1) / BK);
  static int pair_min_k = -1;                  // tuning switch VC_GEMM_PAIR_MINK: shortest reduction (in 64-wide k-blocks) routed to CTA pairs
  if (pair_min_k < 0) { const char* e = getenv("VC_GEMM_PAIR_MINK"); pair_min_k = e ? atoi(e) : 5; }
  // default rule: pairs once every SM pair has work; a wave-tuned choice (above) decides by its cost model instead
  const bool use_pair = tuned == 1 ? true : tuned == 2 ? false :
                        (pair_mode() && (BN == 128 || BN == 160 || BN == 256) && d.N % BN == 0 && k_iters >= pair_min_k &&
                         (m_tiles / 2) * p.n_tiles >= sm_count());

  // A: (K, X, Y, Z) with row pitch lda
  {
    uint64_t dims[4] = {(uint64_t)d.K1, (uint64_t)d.X, (uint64_t)d.Y, (uint64_t)d.Z};
    uint64_t str[3] = {(uint64_t)d.lda * 2, (uint64_t)d.lda * 2 * d.X, (uint64_t)d.lda * 2 * d.X * d.Y};
    uint32_t box[4] = {(uint32_t)BK, (uint32_t)d.bx, (uint32_t)d.by, 1};
    int rc = encode_tmap_f16(&p.tmap_a, d.a, 4, dims, str, box);
    if (rc) return rc;
    if (d.K1 != d.K) {
      uint64_t dims2[4] = {(uint64_t)(d.K - d.K1), (uint64_t)d.X, (uint64_t)d.Y, (uint64_t)d.Z};
      uint64_t str2[3] = {(uint64_t)d.lda2 * 2, (uint64_t)d.lda2 * 2 * d.X, (uint64_t)d.lda2 * 2 * d.X * d.Y};
      rc = encode_tmap_f16(&p.tmap_a2, d.a2, 4, dims2, str2, box);
      if (rc) return rc;
    } else {
      p.tmap_a2 = p.tmap_a;
    }
  }
  {
    uint64_t dims[2] = {(uint64_t)d.K, (uint64_t)d.num_taps * d.N};
    uint64_t str[1] = {(uint64_t)(d.ldw > 0 ? d.ldw : d.K) * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)(use_pair ? BN / 2 : BN)};
    int rc = encode_tmap_f16(&p.tmap_b, d.w, 2, dims, str, box);
    if (rc) return rc;
  }
  p.N = d.N; p.K = d.K; p.K1 = d.K1;
  p.num_taps = d.num_taps;
  for (int t = 0; t < d.num_taps; ++t) { p.tap_dx[t] = d.tap_dx[t]; p.tap_dy[t] = d.tap_dy[t]; }
  p.out = d.out; p.out_f32 = d.out_f32; p.ldo = d.ldo;
  p.bias = d.bias; p.bias_z_div = d.bias_z_div;
  p.res = d.res; p.ldr = d.ldr;
  p.geglu = d.geglu;
  VC_REQUIRE((d.ln_stats == nullptr) == (d.ln_colsum == nullptr), "gemm_tap: ln_stats and ln_colsum go together");
  VC_REQUIRE(!d.ln_stats || (d.num_taps == 1 && d.N % 32 == 0 && d.Y == 1 && d.Z == 1 && !d.a2),
             "gemm_tap: folded LayerNorm needs a plain [M,K] x [N,K] GEMM with N %% 32 == 0");
  VC_REQUIRE(!d.ln_stats || ((reinterpret_cast<uintptr_t>(d.ln_stats) & 7) == 0 && (reinterpret_cast<uintptr_t>(d.ln_colsum) & 15) == 0),
             "gemm_tap: ln_stats / ln_colsum misaligned");
  p.ln_stats = d.ln_stats; p.ln_colsum = d.ln_colsum;
  VC_REQUIRE(!d.ln_part || (d.num_taps == 1 && d.N % 32 == 0 && d.Y == 1 && d.Z == 1 && !d.geglu && d.out && !d.out_f32 &&
                            (reinterpret_cast<uintptr_t>(d.ln_part) & 7) == 0),
             "gemm_tap: LayerNorm partial sums need a plain fp16 [M,N] output with N %% 32 == 0");
  p.ln_part = reinterpret_cast<float2*>(d.ln_part); p.ln_rows = d.X;
  VC_REQUIRE(!d.gn_part || ((d.gn_sub == 10 || d.gn_sub == 8) && d.N % 32 == 0 && d.N % d.gn_sub == 0 && !d.geglu && d.out && !d.out_f32 &&
                            (reinterpret_cast<uintptr_t>(d.gn_part) & 7) == 0),
             "gemm_tap: GroupNorm partial sums need an fp16 output with N %% 32 == 0 and N %% gn_sub == 0 (gn_sub 10 or 8)");
  p.gn_part = reinterpret_cast<float2*>(d.gn_part); p.gn_hp = d.gn_sub / 2; p.gn_nchunks = d.N / 32;
  // 256-bit epilogue accesses need 32-byte aligned rows (true for every activation on the U-Net / VAE path); anything
  // else (odd pitches, the 4- and 3-channel output convs) takes the predicated scalar path inside the kernel
  const bool o_al = ((reinterpret_cast<uintptr_t>(optr) & 31) == 0) && ((long long)d.ldo * esz) % 32 == 0;
  const bool r_al = !d.res || (((reinterpret_cast<uintptr_t>(d.res) & 31) == 0) && ((long long)d.ldr * 2) % 32 == 0);
  p.vec_ok = (o_al && r_al) ? 1 : 0;
  {
    // TMA-store epilogue: fp16 output whose width is whole 32-column chunks and whose rows are 16-byte aligned
    static int tma_store = -1;                 // tuning switch VC_GEMM_TMA_STORE=0 keeps the direct-store epilogue
    if (tma_store < 0) { const char* e = getenv("VC_GEMM_TMA_STORE"); tma_store = (e && e[0] == '0') ? 0 : 1; }
    const int n_out = d.geglu ? d.N / 2 : d.N;
    p.out_tma = (tma_store && d.out && !d.out_f32 && n_out % 32 == 0 && (reinterpret_cast<uintptr_t>(d.out) & 15) == 0 && d.ldo % 8 == 0) ? 1 : 0;
    if (p.out_tma) {
      const uint32_t bw = d.bx < 32 ? d.bx : 32;
      uint64_t dims[4] = {(uint64_t)n_out, (uint64_t)d.X, (uint64_t)d.Y, (uint64_t)d.Z};
      const long long py = d.ldo_y > 0 ? d.ldo_y : (long long)d.ldo * d.X;
      const long long pz = d.ldo_z > 0 ? d.ldo_z : py * d.Y;
      uint64_t str[3] = {(uint64_t)d.ldo * 2, (uint64_t)py * 2, (uint64_t)pz * 2};
      uint32_t box[4] = {32, bw, 32 / bw, 1};
      int rc = encode_tmap_f16(&p.tmap_out, d.out, 4, dims, str, box, 64);
      if (rc) return rc;
    }
    VC_REQUIRE((d.ldo_y == 0 && d.ldo_z == 0) || (p.out_tma && !d.res && !d.ln_part),
               "gemm_tap: strided (ldo_y / ldo_z) outputs need the TMA-store epilogue (fp16, N %% 32 == 0) and no residual");
  }
  if (d.peer && d.peer->mode) {
    // layout switch fused into the epilogue: per-rank destination maps (gemm_common.cuh: GemmPeer)
    const GemmPeerDesc& q = *d.peer;
    VC_REQUIRE(q.mode == 1 || q.mode == 2, "gemm_tap: peer mode %d", q.mode);
    VC_REQUIRE(q.world >= 2 && q.world <= GEMM_PEER_MAX && q.rank >= 0 && q.rank < q.world, "gemm_tap: peer scatter supports 2..%d ranks", GEMM_PEER_MAX);
    VC_REQUIRE(p.out_tma && !d.geglu && !d.ln_part && d.ldo_y == 0 && d.ldo_z == 0, "gemm_tap: peer scatter needs the fp16 TMA-store epilogue");
    VC_REQUIRE(q.HW % q.world == 0 && q.f0[0] == 0 && q.f0[q.world] == q.T && q.B >= 1, "gemm_tap: peer scatter: bad frame / site split");
    GemmPeer& g = p.peer;
    g.mode = q.mode; g.P = q.world; g.me = q.rank;
    g.HW = q.HW; g.HWl = q.HW / q.world; g.T = q.T;
    g.Tl_me = q.f0[q.rank + 1] - q.f0[q.rank];
    for (int r = 0; r <= q.world; ++r) g.f0[r] = q.f0[r];
    const long long rows = (long long)d.X * d.Y * d.Z;
    g.rps = q.mode == 1 ? q.HW : q.T * g.HWl;
    VC_REQUIRE(rows == (q.mode == 1 ? (long long)q.B * g.Tl_me * q.HW : (long long)q.B * q.T * g.HWl), "gemm_tap: peer scatter: %lld rows do not match the layout", rows);
    // every 32-row patch of the epilogue must be a run of consecutive rows of its slab
    const bool row_major_patches = d.Y == 1 || d.bx == d.X || (d.by == 1 && d.X % 32 == 0);
    VC_REQUIRE(row_major_patches && (d.bx >= 32 || 32 % d.bx == 0), "gemm_tap: peer scatter: tile box %dx%d of a %dx%d image is not row-contiguous", d.bx, d.by, d.X, d.Y);
    if (d.Y == 1 && d.Z == 1) { g.wrap = 1; }
    else { g.wrap = 0; VC_REQUIRE((long long)d.X * d.Y == g.rps, "gemm_tap: peer scatter: slab of %d rows expected, tile geometry has %lld", g.rps, (long long)d.X * d.Y); }
    g.div_hwl = make_fastdiv(g.HWl); g.div_rps = make_fastdiv(g.rps); g.div_tl = make_fastdiv(g.Tl_me > 0 ? g.Tl_me : 1);
    for (int r = 0; r < q.world; ++r) {
      VC_REQUIRE(q.dst[r] && (reinterpret_cast<uintptr_t>(q.dst[r]) & 15) == 0, "gemm_tap: peer scatter: destination %d missing / misaligned", r);
      const int tl_r = q.f0[r + 1] - q.f0[r];
      const __half* base = reinterpret_cast<const __half*>(q.dst[r]) + (q.mode == 2 ? (long long)q.rank * g.HWl * d.N : 0);
      uint64_t dims[3] = {(uint64_t)d.N, (uint64_t)g.HWl, (uint64_t)(q.mode == 1 ? (long long)q.B * q.T : (long long)q.B * tl_r)};
      if (dims[2] == 0) dims[2] = 1;              // a rank without frames: nothing is ever routed to it
      uint64_t str[2] = {(uint64_t)d.N * 2, (uint64_t)(q.mode == 1 ? g.HWl : q.HW) * d.N * 2};
      uint32_t box[3] = {32, 32, 1};
      int rc = encode_tmap_f16(&g.map[r], base, 3, dims, str, box, 64);
      if (rc) return rc;
    }
  }
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("VC_GEMM_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    p.debug = dbg;
  }
  const long long total = (use_pair ? (m_tiles + 1) / 2 : m_tiles) * p.n_tiles;
  VC_REQUIRE(total > 0 && total < (1ll << 31), "gemm_tap: tile count %lld out of range", total);
  p.total_tiles = (int)total;
  if (use_pair) return launch_gemm_pair(BN, p, stream);
  switch (BN) {
    case 32: return launch_gemm<32>(p, stream);
    case 64: return launch_gemm<64>(p, stream);
    case 96: return launch_gemm<96>(p, stream);
    case 128: return launch_gemm<128>(p, stream);
    case 160: return launch_gemm<160>(p, stream);
    case 256: return launch_gemm<256>(p, stream);
  }
  set_error("gemm_tap: no kernel for BN=%d", BN);
  return VC_ERR_UNSUPPORTED;
}

}  // namespace vc
