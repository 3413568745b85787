// 64-key-tile variant of flash_attn_d64_kernel that fits THREE CTAs per SM instead of two.  MEASURED on B200 (round 2,
// profiles/r02_ab_micro.txt): it wins where key sequences are short -- cross-attention with 77 / 256 keys 179.5 -> 147.1 us /
// 208.6 -> 186.4 us, self-attention with 576 keys 103 -> 87 us -- and ties / loses from 2304 keys up (0.466 vs 0.464 ms; 3.198 vs
// 3.115 ms at 9216 keys), so flash_attn_d64() dispatches Nk <= 1024 here and keeps the 128-key tiles above.
//
// Why (profiles/README.md, ncu --set full of flash_attn_d64_kernel): the shipped kernel is bound by instruction issue
// of its softmax warps, not by a pipe -- issue slots 65 % busy, MUFU 55 %, tensor 36 %; each softmax warp issues only
// 40 % of its math cycles (fixed-latency dependencies) and a scheduler hosts just two of them (one per co-resident
// CTA).  A third instruction stream per scheduler is what the numbers ask for.  What stops a third CTA is TMEM
// (2 x 256 of 512 columns), registers (168 per thread for a 128-wide score row) and shared memory (2 x 81 KB).
// With 64 keys per tile:
//     TMEM   S 64 + O 64 fp32 columns in one 128-column allocation, P 32 columns (64 fp16) in a second 32-column
//            allocation -> 160 columns per CTA, 480 of 512 for three CTAs
//     regs   64 scores per thread -> __launch_bounds__(192, 3) = 112 registers
//     smem   Q 16 KB + 2 stages x (K 8 KB + V 8 KB) = 48 KB per CTA
// Everything else is the shipped design: SS MMA for S = Q K^T, exp2-domain online softmax with lazy rescale and the
// 25 % polynomial exp2 share, P back to TMEM as fp16, TS MMA for O += P V, separate K / V TMA rings.
// Costs to measure: twice as many tiles (barrier round trips per key halve in size), N = 64 MMAs.
//
// Selection: flash_attn_d64() forwards here for Nk <= 1024; VC_ATTN_BN64=1 / 0 forces / forbids (tools/ab_micro.py prints both).
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace vc {

struct AttnParams64 {
  CUtensorMap tmap_q, tmap_k, tmap_v;
  __half* out;
  int ldo;
  int Nq, Nk;
  int kv_shared;
  float scale_log2;
  int accumulate;
};

static constexpr int A64_BM = 128, A64_BN = 64, A64_D = 64;
static constexpr int A64_Q_BYTES = A64_BM * A64_D * 2;                 // 16 KB
static constexpr int A64_KV_BYTES = A64_BN * A64_D * 2;                // 8 KB
static constexpr int A64_SMEM = A64_Q_BYTES + 4 * A64_KV_BYTES + 1024 + 256;
static constexpr float A64_LAZY = 8.0f;
static constexpr int A64_POLY_PERIOD = 4;

__device__ __forceinline__ float a64_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float a64_ex2_poly(float x) {       // same polynomial as attention.cu: rel. error 1.0e-4
  x = fmaxf(x, -125.0f);
  const float xf = x + 12582912.0f;
  const float f = x - (xf - 12582912.0f);
  float p = fmaf(f, 0.05592204f, 0.24264008f);
  p = fmaf(p, f, 0.69312102f);
  p = fmaf(p, f, 0.99992448f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xf) << 23));
}
__device__ __forceinline__ float a64_ex2_sel(float x, int e) {
  if (A64_POLY_PERIOD > 0 && (e % (2 * A64_POLY_PERIOD)) < 2) return a64_ex2_poly(x);
  return a64_ex2(x);
}

#ifndef VC_ATT_F32X2
#define VC_ATT_F32X2 1
#endif
// two exponentials at once in packed fp32x2 (same arithmetic as attention.cu: ex2_pair)
__device__ __forceinline__ float2 a64_ex2_pair(float2 x, int e) {
  if (A64_POLY_PERIOD > 0 && (e % (2 * A64_POLY_PERIOD)) < 2) {
    x.x = fmaxf(x.x, -125.0f);
    x.y = fmaxf(x.y, -125.0f);
    const float2 xf = __fadd2_rn(x, make_float2(12582912.0f, 12582912.0f));
    const float2 t = __fadd2_rn(xf, make_float2(-12582912.0f, -12582912.0f));
    const float2 f = __ffma2_rn(t, make_float2(-1.f, -1.f), x);
    float2 p = __ffma2_rn(f, make_float2(0.05592204f, 0.05592204f), make_float2(0.24264008f, 0.24264008f));
    p = __ffma2_rn(p, f, make_float2(0.69312102f, 0.69312102f));
    p = __ffma2_rn(p, f, make_float2(0.99992448f, 0.99992448f));
    float2 r;
    r.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(xf.x) << 23));
    r.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(xf.y) << 23));
    return r;
  }
  return make_float2(a64_ex2(x.x), a64_ex2(x.y));
}

__global__ void __launch_bounds__(192, 3) flash_attn_d64_bn64_kernel(const __grid_constant__ AttnParams64 p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + A64_Q_BYTES;                                   // stage s: K at s*16K, V at s*16K + 8K
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A64_Q_BYTES + 4 * A64_KV_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;     // [2]
  uint64_t* k_empty = bars + 3;    // [2]  Q K^T of the stage retired
  uint64_t* s_full = bars + 5;
  uint64_t* s_free = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* o_done = bars + 8;
  uint64_t* v_full = bars + 9;     // [2]
  uint64_t* v_empty = bars + 11;   // [2]  P V of the stage retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);        // [0]: 128-column block (S | O), [1]: 32-column block (P)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * A64_BM, head = blockIdx.y, b = blockIdx.z;
  const int bk = p.kv_shared ? 0 : b;
  const int ntiles = (p.Nk + A64_BN - 1) / A64_BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_q);
    tma_prefetch_desc(&p.tmap_k);
    tma_prefetch_desc(&p.tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);
    mbar_init(p_full, 4);
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_slot[0], 128);          // three co-resident CTAs: 3 x (128 + 32) = 480 of the SM's 512 columns
    tmem_alloc(&tmem_slot[1], 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_a = tmem_slot[0], tmem_b = tmem_slot[1];
  const uint32_t tS = tmem_a, tO = tmem_a + 64, tP = tmem_b;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, A64_Q_BYTES);
      tma_load_4d(sQ, &p.tmap_q, q_full, 0, head, q0, b);
    }
    __syncwarp();
    for (int j = 0; j < ntiles; ++j) {
      const int s = j & 1;
      uint8_t* sk = sKV + s * 2 * A64_KV_BYTES;
      if (j >= 2) mbar_wait(&k_empty[s], ((j >> 1) - 1) & 1);
      if (elect_one()) {
        mbar_expect_tx(&k_full[s], A64_KV_BYTES);
        tma_load_4d(sk, &p.tmap_k, &k_full[s], 0, head, j * A64_BN, bk);
      }
      __syncwarp();
      if (j >= 2) mbar_wait(&v_empty[s], ((j >> 1) - 1) & 1);
      if (elect_one()) {
        mbar_expect_tx(&v_full[s], A64_KV_BYTES);
        tma_load_4d(sk + A64_KV_BYTES, &p.tmap_v, &v_full[s], 0, head, j * A64_BN, bk);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_qk = umma_idesc_f16(128, 64, 0, 0);
    constexpr uint32_t idesc_pv = umma_idesc_f16(128, 64, 0, 1);      // B = V is MN-major
    const uint32_t aQ = smem_u32(sQ);
    auto issue_qk = [&](int j) {
      const int s = j & 1;
      mbar_wait(&k_full[s], (j >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t aK = smem_u32(sKV + s * 2 * A64_KV_BYTES);
#pragma unroll
        for (int k = 0; k < A64_D / 16; ++k)
          umma_ss(tS, umma_desc_sw128(aQ + k * 32), umma_desc_sw128(aK + k * 32), idesc_qk, k > 0 ? 1u : 0u);
        umma_commit(s_full);
        umma_commit(&k_empty[s]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_qk(0);
    for (int j = 0; j < ntiles; ++j) {
      if (j + 1 < ntiles) {
        mbar_wait(s_free, j & 1);
        issue_qk(j + 1);
      }
      mbar_wait(&v_full[j & 1], (j >> 1) & 1);
      mbar_wait(p_full, j & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t aV = smem_u32(sKV + (j & 1) * 2 * A64_KV_BYTES) + A64_KV_BYTES;
#pragma unroll
        for (int k = 0; k < A64_BN / 16; ++k)                          // 16 keys per MMA: 8 TMEM columns of P, 16 rows x 128 B of V
          umma_ts(tO, tP + k * 8, umma_desc_sw128(aV + k * 2048), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(&v_empty[j & 1]);
        umma_commit(o_done);
      }
      __syncwarp();
    }
  } else {
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const float sl2 = p.scale_log2;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      const int valid = min(A64_BN, p.Nk - j * A64_BN);
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld32(tS + lane_off, s0);
      tmem_ld32(tS + lane_off + 32, s1);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(s_free);

      if (valid < A64_BN) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          if (e >= valid) s0[e] = 0xff800000u;                         // -inf: masked keys get probability 0
          if (32 + e >= valid) s1[e] = 0xff800000u;
        }
      }
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        m0 = fmaxf(m0, fmaxf(__uint_as_float(s0[e]), __uint_as_float(s0[e + 1])));
        m1 = fmaxf(m1, fmaxf(__uint_as_float(s0[e + 2]), __uint_as_float(s0[e + 3])));
        m2 = fmaxf(m2, fmaxf(__uint_as_float(s1[e]), __uint_as_float(s1[e + 1])));
        m3 = fmaxf(m3, fmaxf(__uint_as_float(s1[e + 2]), __uint_as_float(s1[e + 3])));
      }
      const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      const float m_cand = fmaxf(m, mx * sl2);
      const bool need = (m_cand - m) > A64_LAZY;                       // j == 0: m = -inf -> true
      float alpha = 1.f;
      if (need) {
        alpha = a64_ex2(m - m_cand);
        l *= alpha;
        m = m_cand;
      }
      const float neg_m = -m;
#if VC_ATT_F32X2
      const float2 sl2v = make_float2(sl2, sl2), negm2 = make_float2(neg_m, neg_m);
      float2 ps0 = make_float2(0.f, 0.f), ps1 = ps0;
      // probabilities, packed in place: s0[0..15] <- s0, s0[16..31] <- s1 (packed fp32x2 arithmetic, see attention.cu)
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        const float2 a = a64_ex2_pair(__ffma2_rn(make_float2(__uint_as_float(s0[e]), __uint_as_float(s0[e + 1])), sl2v, negm2), e);
        ps0 = __fadd2_rn(ps0, a);
        s0[e / 2] = pack_half2(a.x, a.y);
      }
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        const float2 a = a64_ex2_pair(__ffma2_rn(make_float2(__uint_as_float(s1[e]), __uint_as_float(s1[e + 1])), sl2v, negm2), e);
        ps1 = __fadd2_rn(ps1, a);
        s0[16 + e / 2] = pack_half2(a.x, a.y);
      }
      const float2 pst = __fadd2_rn(ps0, ps1);
      l += pst.x + pst.y;
#else
      float ps0 = 0.f, ps1 = 0.f;
      // probabilities, packed in place: s0[0..15] <- s0, s0[16..31] <- s1
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        const float a0 = a64_ex2_sel(fmaf(__uint_as_float(s0[e]), sl2, neg_m), e), a1 = a64_ex2_sel(fmaf(__uint_as_float(s0[e + 1]), sl2, neg_m), e + 1);
        ps0 += a0 + a1;
        s0[e / 2] = pack_half2(a0, a1);
      }
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        const float a0 = a64_ex2_sel(fmaf(__uint_as_float(s1[e]), sl2, neg_m), e), a1 = a64_ex2_sel(fmaf(__uint_as_float(s1[e + 1]), sl2, neg_m), e + 1);
        ps1 += a0 + a1;
        s0[16 + e / 2] = pack_half2(a0, a1);
      }
      l += ps0 + ps1;
#endif
      if (j > 0) {
        mbar_wait(o_done, (j - 1) & 1);                                // P V of the previous tile retired: P and O may be touched
        tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld32(tO + lane_off + c * 32, v);
            tc_wait_ld();
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
            tmem_st32(tO + lane_off + c * 32, v);
          }
        }
      }
      tmem_st32(tP + lane_off, s0);
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(p_full);
    }
    // epilogue
    mbar_wait(o_done, (ntiles - 1) & 1);
    tc_fence_after();
    const float inv = 1.f / l;
    const int row = q0 + r;
    __half* op = p.out + ((long long)b * p.Nq + row) * p.ldo + head * A64_D;
    uint32_t v0[32], v1[32];
    tmem_ld32(tO + lane_off, v0);
    tmem_ld32(tO + lane_off + 32, v1);
    tc_wait_ld();
    if (row < p.Nq) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(g < 4 ? v0[g * 8 + e] : v1[(g - 4) * 8 + e]) * inv;
        uint4* dst = reinterpret_cast<uint4*>(op + g * 8);
        if (p.accumulate) {
          const uint4 u = *dst;
          const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 t = __half22float2(h[e]);
            f[2 * e] += t.x; f[2 * e + 1] += t.y;
          }
        }
        uint4 o;
        o.x = pack_half2(f[0], f[1]); o.y = pack_half2(f[2], f[3]);
        o.z = pack_half2(f[4], f[5]); o.w = pack_half2(f[6], f[7]);
        *dst = o;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_a, 128);
    tmem_dealloc(tmem_b, 32);
  }
}

int flash_attn_bn64_mode() {
  static int mode = -2;
  if (mode == -2) { const char* e = getenv("VC_ATTN_BN64"); mode = !e ? -1 : (e[0] == '1' ? 1 : 0); }
  return mode;
}

int flash_attn_d64_bn64(const AttnDesc& d, cudaStream_t stream) {
  VC_REQUIRE(d.q && d.k && d.v && d.out, "flash_attn: null pointer");
  VC_REQUIRE(d.Nq > 0 && d.Nk > 0 && d.B > 0 && d.heads > 0, "flash_attn: empty problem");
  VC_REQUIRE(d.ldq % 8 == 0 && d.ldk % 8 == 0 && d.ldv % 8 == 0 && d.ldo % 8 == 0, "flash_attn: pitches must be multiples of 8");
  VC_REQUIRE(d.kv_batch_stride % 8 == 0, "flash_attn: kv batch stride must be a multiple of 8");
  AttnParams64 p;
  memset(&p, 0, sizeof(p));
  {
    uint32_t box[4] = {64, 1, 128, 1};
    uint64_t dims[4] = {64, (uint64_t)d.heads, (uint64_t)d.Nq, (uint64_t)d.B};
    uint64_t str[3] = {128, (uint64_t)d.ldq * 2, (uint64_t)d.ldq * 2 * d.Nq};
    int rc = encode_tmap_f16(&p.tmap_q, d.q, 4, dims, str, box);
    if (rc) return rc;
  }
  const int shared = d.kv_batch_stride == 0;
  {
    uint32_t box[4] = {64, 1, 64, 1};                                  // 64 keys per tile
    uint64_t dims[4] = {64, (uint64_t)d.heads, (uint64_t)d.Nk, (uint64_t)(shared ? 1 : d.B)};
    uint64_t strk[3] = {128, (uint64_t)d.ldk * 2, (uint64_t)(shared ? (long long)d.ldk * d.Nk : d.kv_batch_stride) * 2};
    uint64_t strv[3] = {128, (uint64_t)d.ldv * 2, (uint64_t)(shared ? (long long)d.ldv * d.Nk : d.kv_batch_stride) * 2};
    int rc = encode_tmap_f16(&p.tmap_k, d.k, 4, dims, strk, box);
    if (rc) return rc;
    rc = encode_tmap_f16(&p.tmap_v, d.v, 4, dims, strv, box);
    if (rc) return rc;
  }
  p.out = d.out; p.ldo = d.ldo; p.Nq = d.Nq; p.Nk = d.Nk; p.kv_shared = shared;
  p.scale_log2 = d.scale * 1.4426950408889634f;
  p.accumulate = d.accumulate;
  static DeviceOnce configured;
  if (device_once_needed(configured)) {
    VC_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_d64_bn64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A64_SMEM));
    device_once_mark(configured);
  }
  dim3 grid((d.Nq + A64_BM - 1) / A64_BM, d.heads, d.B);
  flash_attn_d64_bn64_kernel<<<grid, 192, A64_SMEM, stream>>>(p);
  VC_CHECK_CUDA(cudaGetLastError());
  return VC_OK;
}

}  // namespace vc
