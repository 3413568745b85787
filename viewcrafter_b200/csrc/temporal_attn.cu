// Temporal self-attention over T <= 32 frames per spatial site (TemporalTransformer attn1/attn2,
// lvdm/modules/attention.py:81-126 with N = T, batch = H*W sites; the reference always takes the naive
// einsum-softmax-einsum path here, attention.py:66).
//
// The problem per (site, head) is a 25x25x64 attention: far below the 128-row granularity of a tcgen05 MMA and
// HBM-bound (3 x T x 128 B in, T x 128 B out per pair), so it runs on warp-level mma.sync m16n8k16 tiles: one warp
// per (site, head), Q/K/V staged in shared memory with 16-byte coalesced loads, S and O accumulators in registers,
// softmax on the accumulator fragments, P re-used in registers as the A operand of P.V (no smem round trip).
// Rows live at (t * sites + site) so no layout transpose of the activations is ever needed.
#include "common.cuh"
#include "kernels.h"

namespace vc {

static constexpr int TA_PITCH = 72;                       // halves per smem row (64 + 8 pad: conflict-free ldmatrix)
static constexpr int TA_WARPS = 4;
static constexpr int TA_SMEM = TA_WARPS * 3 * 32 * TA_PITCH * 2;

__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(TA_WARPS * 32) temporal_attn_kernel(const __half* __restrict__ q, const __half* __restrict__ k,
                                                                      const __half* __restrict__ v, int ld, __half* __restrict__ out,
                                                                      int ldo, int T, long long sites, int heads, float scale_log2) {
  extern __shared__ __align__(16) __half ta_smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __half* Qs = ta_smem + w * 3 * 32 * TA_PITCH;
  __half* Ks = Qs + 32 * TA_PITCH;
  __half* Vs = Ks + 32 * TA_PITCH;
  const int g = lane >> 2, tg = lane & 3;
  const int lrow = lane >> 3, lchunk = (lane & 7) * 8;       // cooperative 16-byte copies: 4 rows x 8 chunks per pass
  const long long pairs = sites * heads;

  for (long long pair = (long long)blockIdx.x * TA_WARPS + w; pair < pairs; pair += (long long)gridDim.x * TA_WARPS) {
    const long long site = pair / heads;
    const int head = (int)(pair % heads);
    __syncwarp();
    // ---- global -> smem (rows >= T are zero) ----
    uint4 rq[8], rk[8], rv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int t = lrow + 4 * i;
      if (t < T) {
        const long long off = ((long long)t * sites + site) * ld + head * 64 + lchunk;
        rq[i] = *reinterpret_cast<const uint4*>(q + off);
        rk[i] = *reinterpret_cast<const uint4*>(k + off);
        rv[i] = *reinterpret_cast<const uint4*>(v + off);
      } else {
        rq[i] = rk[i] = rv[i] = make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int t = lrow + 4 * i;
      *reinterpret_cast<uint4*>(Qs + t * TA_PITCH + lchunk) = rq[i];
      *reinterpret_cast<uint4*>(Ks + t * TA_PITCH + lchunk) = rk[i];
      *reinterpret_cast<uint4*>(Vs + t * TA_PITCH + lchunk) = rv[i];
    }
    __syncwarp();

    // ---- S = Q K^T : [32 x 32], k = 64 ----
    float s[2][4][4];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[mi][ni][e] = 0.f;
    const int mat = lane >> 3, mr = lane & 7;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[2][4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
        ldsm_x4(a[mi][0], a[mi][1], a[mi][2], a[mi][3], Qs + (mi * 16 + (mat & 1) * 8 + mr) * TA_PITCH + kk * 16 + (mat >> 1) * 8);
#pragma unroll
      for (int np = 0; np < 2; ++np) {                       // two n-tiles (16 keys) per ldmatrix.x4
        uint32_t b0, b1, b2, b3;
        ldsm_x4(b0, b1, b2, b3, Ks + (np * 16 + (mat >> 1) * 8 + mr) * TA_PITCH + kk * 16 + (mat & 1) * 8);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
          mma16816(s[mi][2 * np], a[mi], b0, b1);
          mma16816(s[mi][2 * np + 1], a[mi], b2, b3);
        }
      }
    }

    // ---- softmax over the key axis (columns); each thread owns rows g / g+8 of both m-tiles ----
    uint32_t p[2][2][4];                                      // P as A fragments: [m-tile][k-step of 16 keys][4]
    float inv_l[2][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {                           // h = 0: row g, h = 1: row g + 8
        float mx = -INFINITY;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = ni * 8 + 2 * tg + e;
            float val = s[mi][ni][2 * h + e];
            if (col >= T) val = -INFINITY;
            s[mi][ni][2 * h + e] = val;
            mx = fmaxf(mx, val);
          }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float off = mx * scale_log2;
        float l = 0.f;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float pv = exp2f(fmaf(s[mi][ni][2 * h + e], scale_log2, -off));
            s[mi][ni][2 * h + e] = pv;
            l += pv;
          }
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        inv_l[mi][h] = 1.f / l;
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        p[mi][kk][0] = pack_half2(s[mi][2 * kk][0], s[mi][2 * kk][1]);
        p[mi][kk][1] = pack_half2(s[mi][2 * kk][2], s[mi][2 * kk][3]);
        p[mi][kk][2] = pack_half2(s[mi][2 * kk + 1][0], s[mi][2 * kk + 1][1]);
        p[mi][kk][3] = pack_half2(s[mi][2 * kk + 1][2], s[mi][2 * kk + 1][3]);
      }
    }

    // ---- O = P V : [32 x 64], k = 32 keys ----
    float o[2][8][4];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 8; ++ni)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[mi][ni][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {                       // two d-tiles (16 columns) per ldmatrix.x4.trans
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(b0, b1, b2, b3, Vs + (kk * 16 + (mat & 1) * 8 + mr) * TA_PITCH + np * 16 + (mat >> 1) * 8);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
          mma16816(o[mi][2 * np], p[mi][kk], b0, b1);
          mma16816(o[mi][2 * np + 1], p[mi][kk], b2, b3);
        }
      }
    }

    // ---- O / l -> smem (reuse the Q tile) -> coalesced 16-byte row stores ----
    __syncwarp();
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) {
        *reinterpret_cast<uint32_t*>(Qs + (mi * 16 + g) * TA_PITCH + ni * 8 + 2 * tg) =
            pack_half2(o[mi][ni][0] * inv_l[mi][0], o[mi][ni][1] * inv_l[mi][0]);
        *reinterpret_cast<uint32_t*>(Qs + (mi * 16 + g + 8) * TA_PITCH + ni * 8 + 2 * tg) =
            pack_half2(o[mi][ni][2] * inv_l[mi][1], o[mi][ni][3] * inv_l[mi][1]);
      }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int t = lrow + 4 * i;
      if (t < T)
        *reinterpret_cast<uint4*>(out + ((long long)t * sites + site) * ldo + head * 64 + lchunk) =
            *reinterpret_cast<const uint4*>(Qs + t * TA_PITCH + lchunk);
    }
  }
}

int temporal_attn(const __half* q, const __half* k, const __half* v, int ld, __half* out, int ldo, int T, long long sites,
                  int heads, float scale, cudaStream_t stream) {
  VC_REQUIRE(q && k && v && out, "temporal_attn: null pointer");
  VC_REQUIRE(T >= 1 && T <= 32, "temporal_attn: T=%d unsupported (1..32)", T);
  VC_REQUIRE(ld % 8 == 0 && ldo % 8 == 0, "temporal_attn: pitches must be multiples of 8");
  static DeviceOnce configured;
  if (device_once_needed(configured)) {
    VC_CHECK_CUDA(cudaFuncSetAttribute(temporal_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM));
    device_once_mark(configured);
  }
  const long long pairs = sites * heads;
  long long blocks = (pairs + TA_WARPS - 1) / TA_WARPS;
  const long long cap = (long long)sm_count() * 4;          // 4 resident blocks per SM (55 KB smem each), grid-stride beyond
  if (blocks > cap) blocks = cap;
  temporal_attn_kernel<<<(unsigned)blocks, TA_WARPS * 32, TA_SMEM, stream>>>(q, k, v, ld, out, ldo, T, sites, heads,
                                                                            scale * 1.4426950408889634f);
  VC_CHECK_CUDA(cudaGetLastError());
  return VC_OK;
}

}  // namespace vc
