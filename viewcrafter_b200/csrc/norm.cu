// HBM-bound normalisation kernels on channels-last fp16 activations.
//   groupnorm_nhwc : GroupNorm(32) (+SiLU) over [samples, rows, C]; C may be the concat of two tensors.
//                    pass 1 = per-CTA partial (sum, sumsq) per group (deterministic, no atomics to global),
//                    pass 2 = finalize stats in smem + normalise/affine/SiLU with 16-byte vector I/O.
//   layernorm_rows : LayerNorm over the last dim, one warp per token row, values held in registers.
// Reference semantics: lvdm/basics.py:76-87 (fp32 GroupNorm), attention.py:265,331 (eps 1e-6),
// openaimodel3d.py:256-265 (5-D GroupNorm in TemporalConvBlock), torch.nn.LayerNorm (eps 1e-5).
#include "common.cuh"
#include "kernels.h"

namespace vc {

static constexpr int GN_MAX_SPLITS = 512;

size_t groupnorm_ws_bytes(int samples) { return (size_t)samples * GN_MAX_SPLITS * 64 * sizeof(float) + (size_t)samples * sizeof(unsigned int); }

struct GnGeom {
  int C, C1, C2, vecs, ppi, cg, splits;
  long long rows, rows_per_split;
  int stat_splits;        // partial-statistics records per sample the apply pass sums
  long long stat_rows;    // rows the statistics cover (== rows, or the global row count when sharded across GPUs)
};

__device__ __forceinline__ uint4 gn_load(const __half* x1, const __half* x2, const GnGeom& g, long long sample, long long row, int c) {
  const long long r = sample * g.rows + row;
  const __half* p = (c < g.C1) ? (x1 + r * g.C1 + c) : (x2 + r * g.C2 + (c - g.C1));
  return *reinterpret_cast<const uint4*>(p);
}

__device__ __forceinline__ void gn_stats_dev(const __half* __restrict__ x1, const __half* __restrict__ x2, const GnGeom& g,
                                                        float* __restrict__ partial) {
  extern __shared__ float red[];   // [2*C]
  const int tid = threadIdx.x;
  const int v = tid % g.vecs, pl = tid / g.vecs;
  const int split = blockIdx.x, sample = blockIdx.y;
  for (int i = tid; i < 2 * g.C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float s[8], ss[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = ss[e] = 0.f;
  const long long r0 = (long long)split * g.rows_per_split;
  const long long r1 = min(g.rows, r0 + g.rows_per_split);
  // 4 independent 16-byte loads in flight per thread (the kernel is pure streaming: latency must be covered by MLP)
  for (long long r = r0 + pl; r < r1; r += 4ll * g.ppi) {
    uint4 u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long ri = r + (long long)i * g.ppi;
      u[i] = ri < r1 ? gn_load(x1, x2, g, sample, ri, v * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __half2* h = reinterpret_cast<const __half2*>(&u[i]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h[e]);
        s[2 * e] += f.x; ss[2 * e] += f.x * f.x;
        s[2 * e + 1] += f.y; ss[2 * e + 1] += f.y * f.y;
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    atomicAdd(&red[v * 8 + e], s[e]);
    atomicAdd(&red[g.C + v * 8 + e], ss[e]);
  }
  __syncthreads();
  if (tid < 64) {
    const int grp = tid >> 1, which = tid & 1;
    float acc = 0.f;
    for (int c = grp * g.cg; c < (grp + 1) * g.cg; ++c) acc += red[which * g.C + c];
    partial[((long long)sample * g.splits + split) * 64 + tid] = acc;
  }
}

__device__ __forceinline__ void gn_apply_dev(const __half* __restrict__ x1, const __half* __restrict__ x2, const GnGeom& g,
                                                        const float* __restrict__ partial, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, int silu, __half* __restrict__ out) {
  __shared__ float mean_s[32], rstd_s[32];
  const int tid = threadIdx.x;
  const int sample = blockIdx.y;
  if (tid < 32) {
    double sum = 0.0, sq = 0.0;
    const float* pp = partial + (long long)sample * g.stat_splits * 64;
    for (int sp = 0; sp < g.stat_splits; ++sp) {
      sum += pp[sp * 64 + tid * 2];
      sq += pp[sp * 64 + tid * 2 + 1];
    }
    const double n = (double)g.stat_rows * g.cg;
    const double m = sum / n;
    double var = sq / n - m * m;
    if (var < 0.0) var = 0.0;
    mean_s[tid] = (float)m;
    rstd_s[tid] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const int v = tid % g.vecs, pl = tid / g.vecs;
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = v * 8 + e;
    const int grp = c / g.cg;
    const float a = rstd_s[grp] * gamma[c];
    sc[e] = a;
    sh[e] = beta[c] - mean_s[grp] * a;
  }
  const long long r0 = (long long)blockIdx.x * g.rows_per_split;
  const long long r1 = min(g.rows, r0 + g.rows_per_split);
  for (long long r = r0 + pl; r < r1; r += 4ll * g.ppi) {
    uint4 u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long ri = r + (long long)i * g.ppi;
      if (ri < r1) u[i] = gn_load(x1, x2, g, sample, ri, v * 8);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long ri = r + (long long)i * g.ppi;
      if (ri >= r1) break;
      const __half2* h = reinterpret_cast<const __half2*>(&u[i]);
      float f[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 t = __half22float2(h[e]);
        f[2 * e] = t.x; f[2 * e + 1] = t.y;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float y = f[e] * sc[e] + sh[e];
        f[e] = silu ? silu_f(y) : y;
      }
      uint4 o;
      o.x = pack_half2(f[0], f[1]); o.y = pack_half2(f[2], f[3]);
      o.z = pack_half2(f[4], f[5]); o.w = pack_half2(f[6], f[7]);
      *reinterpret_cast<uint4*>(out + ((long long)sample * g.rows + ri) * g.C + v * 8) = o;
    }
  }
}

__global__ void __launch_bounds__(1024) gn_stats_kernel(const __half* __restrict__ x1, const __half* __restrict__ x2, GnGeom g,
                                                        float* __restrict__ partial) {
  gn_stats_dev(x1, x2, g, partial);
}
__global__ void __launch_bounds__(1024) gn_apply_kernel(const __half* __restrict__ x1, const __half* __restrict__ x2, GnGeom g,
                                                        const float* __restrict__ partial, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, int silu, __half* __restrict__ out) {
  gn_apply_dev(x1, x2, g, partial, gamma, beta, eps, silu, out);
}
// Fused single launch: statistics pass, a grid-wide rendezvous of the CTAs of one sample (all CTAs are co-resident by
// construction -- the host checks the occupancy), then the normalise pass, whose re-read of x is served by the 126 MB L2
// for everything but the largest 5-D tensors: HBM traffic drops from 3 passes to ~2.
__global__ void __launch_bounds__(1024) gn_fused_kernel(const __half* __restrict__ x1, const __half* __restrict__ x2, GnGeom g,
                                                        float* __restrict__ partial, unsigned int* __restrict__ counters,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int silu,
                                                        __half* __restrict__ out) {
  gn_stats_dev(x1, x2, g, partial);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(&counters[blockIdx.y], 1u);
    while (*reinterpret_cast<volatile unsigned int*>(&counters[blockIdx.y]) < (unsigned)g.splits) __nanosleep(40);
    __threadfence();
  }
  __syncthreads();
  gn_apply_dev(x1, x2, g, partial, gamma, beta, eps, silu, out);
}

static int gn_geometry(GnGeom& g, const __half* x1, int C1, const __half* x2, int C2, int samples, long long rows_per_sample) {
  const int C = C1 + (x2 ? C2 : 0);
  VC_REQUIRE(x1, "groupnorm: null pointer");
  VC_REQUIRE(C % 32 == 0 && C1 % 8 == 0 && (!x2 || C2 % 8 == 0) && C <= 8192, "groupnorm: unsupported channels C1=%d C2=%d", C1, C2);
  VC_REQUIRE(samples >= 1 && rows_per_sample >= 1, "groupnorm: empty input");
  g.C = C; g.C1 = C1; g.C2 = x2 ? C2 : 0;
  g.vecs = C / 8;
  g.ppi = 512 / g.vecs > 0 ? 512 / g.vecs : 1;
  g.cg = C / 32;
  g.rows = rows_per_sample;
  int splits = (2 * sm_count() + samples - 1) / samples;
  const long long max_useful = (rows_per_sample + g.ppi - 1) / g.ppi;
  if (splits > max_useful) splits = (int)max_useful;
  if (splits > GN_MAX_SPLITS) splits = GN_MAX_SPLITS;
  if (splits < 1) splits = 1;
  g.splits = splits;
  g.rows_per_split = (rows_per_sample + splits - 1) / splits;
  g.stat_splits = splits;
  g.stat_rows = rows_per_sample;
  return VC_OK;
}

int groupnorm_nhwc(const __half* x1, int C1, const __half* x2, int C2, int samples, long long rows_per_sample,
                   const float* gamma, const float* beta, float eps, int silu, __half* out, float* partial_ws,
                   size_t ws_bytes, cudaStream_t stream) {
  VC_REQUIRE(out && gamma && beta && partial_ws, "groupnorm: null pointer");
  GnGeom g;
  int rc = gn_geometry(g, x1, C1, x2, C2, samples, rows_per_sample);
  if (rc) return rc;
  VC_REQUIRE(ws_bytes >= (size_t)samples * g.splits * 64 * sizeof(float), "groupnorm: workspace too small");
  const int threads = g.vecs * g.ppi;
  dim3 grid(g.splits, samples);
  const size_t smem = 2 * g.C * sizeof(float);
  // fused path: needs every CTA resident at once (spin rendezvous) and room for the per-sample counters after the partials
  const size_t part_bytes = (size_t)samples * g.splits * 64 * sizeof(float);
  int per_sm = 0;
  VC_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gn_fused_kernel, threads, smem));
  const long long capacity = (long long)per_sm * sm_count();
  if (capacity < (long long)g.splits * samples && capacity >= samples) {   // shrink the split so that every CTA is resident
    g.splits = (int)(capacity / samples);
    g.rows_per_split = (rows_per_sample + g.splits - 1) / g.splits;
    g.stat_splits = g.splits;
    grid = dim3(g.splits, samples);
  }
  const bool fused = capacity >= (long long)g.splits * samples && ws_bytes >= part_bytes + samples * sizeof(unsigned int);
  if (fused) {
    unsigned int* counters = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(partial_ws) + part_bytes);
    VC_CHECK_CUDA(cudaMemsetAsync(counters, 0, samples * sizeof(unsigned int), stream));
    gn_fused_kernel<<<grid, threads, smem, stream>>>(x1, x2, g, partial_ws, counters, gamma, beta, eps, silu, out);
    VC_CHECK_CUDA(cudaGetLastError());
    return VC_OK;
  }
  gn_stats_kernel<<<grid, threads, smem, stream>>>(x1, x2, g, partial_ws);
  VC_CHECK_CUDA(cudaGetLastError());
  gn_apply_kernel<<<grid, threads, 0, stream>>>(x1, x2, g, partial_ws, gamma, beta, eps, silu, out);
  VC_CHECK_CUDA(cudaGetLastError());
  return VC_OK;
}

// Split form for statistics that span several GPUs (site-sharded 5-D GroupNorm): pass 1 leaves (sum, sumsq) per group in
// stats[samples][32][2]; the caller all-reduces that tiny buffer; pass 2 normalises with the global row count.
__global__ void gn_finalize_kernel(const float* __restrict__ partial, int splits, float* __restrict__ stats) {
  const int sample = blockIdx.x, t = threadIdx.x;   // 64 threads
  float acc = 0.f;
  for (int sp = 0; sp < splits; ++sp) acc += partial[((long long)sample * splits + sp) * 64 + t];
  stats[sample * 64 + t] = acc;
}

int groupnorm_stats(const __half* x1, int C1, const __half* x2, int C2, int samples, long long rows_per_sample, float* stats,
                    float* partial_ws, size_t ws_bytes, cudaStream_t stream) {
  VC_REQUIRE(stats && partial_ws, "groupnorm_stats: null pointer");
  GnGeom g;
  int rc = gn_geometry(g, x1, C1, x2, C2, samples, rows_per_sample);
  if (rc) return rc;
  VC_REQUIRE(ws_bytes >= (size_t)samples * g.splits * 64 * sizeof(float), "groupnorm: workspace too small");
  dim3 grid(g.splits, samples);
  gn_stats_kernel<<<grid, g.vecs * g.ppi, 2 * g.C * sizeof(float), stream>>>(x1, x2, g, partial_ws);
  VC_CHECK_CUDA(cudaGetLastError());
  gn_finalize_kernel<<<samples, 64, 0, stream>>>(partial_ws, g.splits, stats);
  VC_CHECK_CUDA(cudaGetLastError());
  return VC_OK;
}

int groupnorm_apply(const __half* x1, int C1, const __half* x2, int C2, int samples, long long rows_per_sample,
                    const float* stats, long long stat_rows, const float* gamma, const float* beta, float eps, int silu, __half* out,
                    cudaStream_t stream) {
  VC_REQUIRE(out && gamma && beta && stats && stat_rows >= rows_per_sample, "groupnorm_apply: bad args");
  GnGeom g;
  int rc = gn_geometry(g, x1, C1, x2, C2, samples, rows_per_sample);
  if (rc) return rc;
  g.stat_splits = 1;
  g.stat_rows = stat_rows;
  dim3 grid(g.splits, samples);
  gn_apply_kernel<<<grid, g.vecs * g.ppi, 0, stream>>>(x1, x2, g, stats, gamma, beta, eps, silu, out);
  VC_CHECK_CUDA(cudaGetLastError());
  return VC_OK;
}

// ------------------------------------------------------------------------------------------------
// Persistent warps: each warp walks rows with a grid stride and keeps TWO rows in flight (all their 16-byte loads are
// issued before the first reduction) so the DRAM latency is covered by memory-level parallelism, not by block churn.
template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, long long rows, int C, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, __half* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int vecs = C >> 3;
  const float invC = 1.f / (float)C;
  for (long long row0 = gw * 2; row0 < rows; row0 += nwarps * 2) {
    uint4 raw[2][MAXV];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const long long row = row0 + rr;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int v = lane + i * 32;
        raw[rr][i] = (row < rows && v < vecs) ? *reinterpret_cast<const uint4*>(x + row * C + v * 8) : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const long long row = row0 + rr;
      if (row >= rows) break;                                   // warp-uniform
      float f[MAXV][8];
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const __half2* h = reinterpret_cast<const __half2*>(&raw[rr][i]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 t = __half22float2(h[e]);
          f[i][2 * e] = t.x; f[i][2 * e + 1] = t.y;
          sum += t.x + t.y;                                     // out-of-range vectors were loaded as zeros
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum * invC;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        if (lane + i * 32 < vecs) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float d = f[i][e] - mean;
            sq += d * d;
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      const float rstd = rsqrtf(sq * invC + eps);
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int v = lane + i * 32;
        if (v < vecs) {
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          float y[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) y[e] = (f[i][e] - mean) * rstd * gg[e] + bb[e];
          uint4 o;
          o.x = pack_half2(y[0], y[1]); o.y = pack_half2(y[2], y[3]);
          o.z = pack_half2(y[4], y[5]); o.w = pack_half2(y[6], y[7]);
          *reinterpret_cast<uint4*>(out + row * C + v * 8) = o;
        }
      }
    }
  }
}

int layernorm_rows(const __half* x, long long rows, int C, const float* gamma, const float* beta, float eps, __half* out,
                   cudaStream_t stream) {
  VC_REQUIRE(x && out && gamma && beta, "layernorm: null pointer");
  VC_REQUIRE(C % 8 == 0 && C <= 2048 && rows > 0, "layernorm: unsupported C=%d rows=%lld", C, rows);
  const int wpb = 8;
  long long blocks = (rows + 2 * wpb - 1) / (2 * wpb);
  const long long cap = (long long)sm_count() * 8;              // 8 x 256 threads = 64 warps per SM
  if (blocks > cap) blocks = cap;
  const int vecs = C / 8;
  if (vecs <= 64)
    layernorm_kernel<2><<<(unsigned)blocks, wpb * 32, 0, stream>>>(x, rows, C, gamma, beta, eps, out);
  else if (vecs <= 160)
    layernorm_kernel<5><<<(unsigned)blocks, wpb * 32, 0, stream>>>(x, rows, C, gamma, beta, eps, out);
  else
    layernorm_kernel<8><<<(unsigned)blocks, wpb * 32, 0, stream>>>(x, rows, C, gamma, beta, eps, out);
  VC_CHECK_CUDA(cudaGetLastError());
  return VC_OK;
}

}  // namespace vc
