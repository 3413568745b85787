// Shared pieces of the tap-GEMM kernels (gemm_tap.cu: one CTA per tile, gemm_tap2.cu: CTA pair per 256-row tile):
// parameter block, tile enumeration and the TMEM -> registers -> global epilogue.
#pragma once
#include "common.cuh"

namespace vc {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int MAX_TAPS = 9;
static constexpr int EPI_WARPS = 16;
static constexpr int EPI_PER_QUAD = EPI_WARPS / 4;
static constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;

struct GemmParams {
  CUtensorMap tmap_a;
  CUtensorMap tmap_a2;
  CUtensorMap tmap_b;
  int tiles_x, tiles_y, Z;
  int bx, by;
  int X, Y;
  int N, K, K1;          // K1 = channels served by tmap_a (K1 == K when single source)
  int num_taps;
  int tap_dx[MAX_TAPS];
  int tap_dy[MAX_TAPS];
  int n_tiles;
  int total_tiles;       // 1-CTA kernel: m_tiles * n_tiles; pair kernel: ceil(m_tiles / 2) * n_tiles
  __half* out;
  float* out_f32;
  int ldo;
  const float* bias;
  int bias_z_div;
  const __half* res;
  int ldr;
  int geglu;
  int vec_ok;            // rows are 32-byte aligned: the 256-bit epilogue path may be used
  int debug;             // profiling aid (VC_GEMM_DEBUG): 1 = skip the MMAs (feed rate only), 2 = skip TMA (MMA rate only),
                         // 4 = skip the epilogue body; results are garbage in these modes
};

struct TileCoord {
  int x0, y0, z;
};

// m-tile index -> tile origin; indices past the last m-tile give z >= Z (TMA zero-fills, the epilogue masks the rows)
__device__ __forceinline__ TileCoord tile_coord_m(const GemmParams& p, int m) {
  TileCoord t;
  const int tx = m % p.tiles_x;
  m /= p.tiles_x;
  const int ty = m % p.tiles_y;
  t.z = m / p.tiles_y;
  t.x0 = tx * p.bx;
  t.y0 = ty * p.by;
  return t;
}

#ifdef __CUDACC__
// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): one instruction moves a full 32-byte sector per thread
__device__ __forceinline__ void st_global_256(void* ptr, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
               "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void ld_global_256(const void* ptr, uint32_t (&v)[8]) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(ptr)
               : "memory");
}

// exact-erf GELU (attention.py:415-422 uses F.gelu), erf by Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7): 2 MUFU + ~12 FMA/ALU
__device__ __forceinline__ float gelu_epilogue(float x) {
  const float z = x * 0.70710678118654752440f;
  const float az = fabsf(z);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, az, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(az * az * -1.4426950408889634f));
  const float erf_abs = fmaf(-poly * t, e, 1.0f);
  const float hx = 0.5f * x;
  return fmaf(hx, copysignf(erf_abs, z), hx);
}

// Epilogue of one 128 x BN accumulator (this CTA's TMEM, column base `tacc`), executed by the 16 epilogue warps
// (warp index 2..17).  Four warps per TMEM lane quadrant take interleaved 32-column chunks; each thread owns one row:
// TMEM -> registers -> (+bias / GEGLU / +residual) -> 256-bit global stores (whole 32-byte sectors).
template <int BN>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmParams& p, const TileCoord& tc, int n_tile, uint32_t tacc, int warp,
                                                   int lane) {
  const int q = warp & 3;                       // TMEM lane quadrant this warp may access
  const int sub = (warp - 2) >> 2;              // which of the EPI_PER_QUAD warps of this quadrant
  constexpr int HALF = BN / 2;
  const int nchunks = p.geglu ? HALF / 32 : BN / 32;
  const int n_out = p.geglu ? p.N / 2 : p.N;
  const int R = q * 32 + lane;                  // accumulator row owned by this thread
  const int x = tc.x0 + (R % p.bx), y = tc.y0 + (R / p.bx);
  const bool row_ok = x < p.X && y < p.Y && tc.z < p.Z;
  const long long orow = ((long long)tc.z * p.Y + y) * p.X + x;
  const int n0 = n_tile * BN;
  const int ocol0 = p.geglu ? n_tile * HALF : n0;
  const float* bias = p.bias ? p.bias + (long long)(p.bias_z_div > 0 ? min(tc.z, p.Z - 1) / p.bias_z_div : 0) * p.N : nullptr;
  const uint32_t trow = tacc + ((uint32_t)(q * 32) << 16);

#pragma unroll 1
  for (int c = sub; c < nchunks; c += EPI_PER_QUAD) {
    float f[32];
    __syncwarp();
    if (!p.geglu) {
      uint32_t v[32];
      tmem_ld32(trow + c * 32, v);
      tc_wait_ld();
      const int nb = n0 + c * 32;
      if (nb >= p.N) break;                      // warp-uniform
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      if (bias) {
        if (nb + 32 <= p.N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + nb + j));
            f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nb + j < p.N) f[j] += __ldg(bias + nb + j);
        }
      }
    } else {
      // GEGLU: tile columns [0,BN/2) are values, [BN/2,BN) the matching gates (weights were interleaved per tile).
      uint32_t a[32], g[32];
      tmem_ld32(trow + c * 32, a);
      tmem_ld32(trow + HALF + c * 32, g);
      tc_wait_ld();
      const int nv = n0 + c * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bg = ba;
        if (bias) {
          ba = __ldg(reinterpret_cast<const float4*>(bias + nv + j));
          bg = __ldg(reinterpret_cast<const float4*>(bias + nv + HALF + j));
        }
        f[j] = (__uint_as_float(a[j]) + ba.x) * gelu_epilogue(__uint_as_float(g[j]) + bg.x);
        f[j + 1] = (__uint_as_float(a[j + 1]) + ba.y) * gelu_epilogue(__uint_as_float(g[j + 1]) + bg.y);
        f[j + 2] = (__uint_as_float(a[j + 2]) + ba.z) * gelu_epilogue(__uint_as_float(g[j + 2]) + bg.z);
        f[j + 3] = (__uint_as_float(a[j + 3]) + ba.w) * gelu_epilogue(__uint_as_float(g[j + 3]) + bg.w);
      }
    }
    const int col0 = ocol0 + c * 32;
    if (!row_ok || col0 >= n_out) continue;
    if (col0 + 32 <= n_out && p.vec_ok) {
      if (p.res) {
        const __half* rp = p.res + orow * p.ldr + col0;      // plain loads: res may alias out (in-place residual)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint32_t u[8];
          ld_global_256(rp + j * 16, u);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&u[e]));
            f[j * 16 + 2 * e] += t.x; f[j * 16 + 2 * e + 1] += t.y;
          }
        }
      }
      if (p.out_f32) {
        float* op = p.out_f32 + orow * p.ldo + col0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t u[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) u[e] = __float_as_uint(f[j * 8 + e]);
          st_global_256(op + j * 8, u);
        }
      } else {
        __half* op = p.out + orow * p.ldo + col0;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint32_t u[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) u[e] = pack_half2(f[j * 16 + 2 * e], f[j * 16 + 2 * e + 1]);
          st_global_256(op + j * 16, u);
        }
      }
    } else {
      // ragged N tail / unaligned pitch (e.g. the 320->4 output conv): predicated scalar path
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        if (col0 + e < n_out) {
          float t = f[e];
          if (p.res) t += __half2float(p.res[orow * p.ldr + col0 + e]);
          if (p.out_f32) p.out_f32[orow * p.ldo + col0 + e] = t;
          else p.out[orow * p.ldo + col0 + e] = __float2half_rn(t);
        }
      }
    }
  }
}
#endif  // __CUDACC__

int launch_gemm_pair(int BN, const GemmParams& p, cudaStream_t stream);

}  // namespace vc
