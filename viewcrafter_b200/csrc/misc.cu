// Small / HBM-bound kernels of the denoise step:
// nearest-2x upsample, stride-2 im2col, layout conversions, the timestep/fps embedding MLP pieces
// and the fused DDIM update.
#include "common.cuh"
#include "kernels.h"

namespace vc {

// ------------------------------------------------------------------------------------------------
// nearest 2x upsample (F.interpolate scale 2, openaimodel3d.py:101-104 / ae_modules.py:123), NHWC, 16-byte vectors
// ------------------------------------------------------------------------------------------------
__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, long long total, int H, int W, int vecs) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    long long p = i / vecs;
    const int xo = (int)(p % (2 * W)); p /= (2 * W);
    const int yo = (int)(p % (2 * H));
    const long long n = p / (2 * H);
    out[i] = x[((n * H + (yo >> 1)) * W + (xo >> 1)) * vecs + v];
  }
}
int upsample2x_nhwc(const __half* x, __half* out, int N, int H, int W, int C, cudaStream_t stream) {
  VC_REQUIRE(x && out && C % 8 == 0, "upsample2x: bad args");
  const long long total = (long long)N * 4 * H * W * (C / 8);
  const int blocks = (int)min((long long)sm_count() * 16, (total + 255) / 256);
  upsample2x_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(out), total, H, W, C / 8);
  VC_CHECK_CUDA(cudaGetLastError());
  return VC_OK;
}

// im2col for the stride-2 3x3 downsample conv (openaimodel3d.py:51-77): out [N*Ho*Wo, 9*C], tap-major then channel.
__global__ void im2col_s2_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, long long total, int H, int W, int vecs,
                                 int pad_lo, int Ho, int Wo) {
  const uint4 zero = make_uint4(0, 0, 0, 0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    long long p = i / vecs;
    const int tap = (int)(p % 9); p /= 9;
    const int xo = (int)(p % Wo); p /= Wo;
    const int yo = (int)(p % Ho);
    const long long n = p / Ho;
    const int yi = 2 * yo - pad_lo + tap / 3, xi = 2 * xo - pad_lo + tap % 3;
    out[i] = (yi >= 0 && yi < H && xi >= 0 && xi < W) ? x[((n * H + yi) * W + xi) * vecs + v] : zero;
  }
}
int im2col3x3_s2_nhwc(const __half* x, __half* out, int N, int H, int W, int C, int pad_lo, int Ho, int Wo, cudaStream_t stream) {
  VC_REQUIRE(x && out && C % 8 == 0, "im2col: bad args");
  const long long total = (long long)N * Ho * Wo * 9 * (C / 8);
  const int blocks = (int)min((long long)sm_count() * 16, (total + 255) / 256);
  im2col_s2_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(out), total, H, W, C / 8,
                                               pad_lo, Ho, Wo);
  VC_CHECK_CUDA(cudaGetLastError());
  return VC_OK;
}

// ------------------------------------------------------------------------------------------------
// layout conversions at the boundary ([B,C,T,H,W] fp32 <-> [(B T) H W, C] fp16/fp32)
// ------------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, __half* __restrict__ out, int B, int C, int T, long long HW,
                                    int c_off, int ldo) {
  const long long total = (long long)B * C * T * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i % HW;
    long long r = i / HW;
    const int t = (int)(r % T); r /= T;
    const int c = (int)(r % C);
    const long long b = r / C;
    out[((b * T + t) * HW + p) * ldo + c_off + c] = __float2half_rn(x[i]);
  }
}
int nchw_to_nhwc_f16(const float* x, __half* out, int B, int C, int T, long long HW, int c_off, int ldo, cudaStream_t stream) {
  VC_REQUIRE(x && out, "nchw_to_nhwc: null pointer");
  const long long total = (long long)B * C * T * HW;
  const int blocks = (int)min((long long)sm_count() * 16, (total + 255) / 256);
  nchw_to_nhwc_kernel<<<blocks, 256, 0, stream>>>(x, out, B, C, T, HW, c_off, ldo);
  VC_CHECK_CUDA(cudaGetLastError());
  return VC_OK;
}

__global__ void nhwc_to_ncthw_kernel(const float* __restrict__ x, int ldx, float* __restrict__ out, int B, int C, int T, long long HW) {
  const long long total = (long long)B * C * T * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i % HW;
    long long r = i / HW;
    const int t = (int)(r % T); r /= T;
    const int c = (int)(r % C);
    const long long b = r / C;
    out[i] = x[((b * T + t) * HW + p) * ldx + c];
  }
}
int nhwc_to_ncthw_f32(const float* x, int ldx, float* out, int B, int C, int T, long long HW, cudaStream_t stream) {
  VC_REQUIRE(x && out, "nhwc_to_ncthw: null pointer");
  const long long total = (long long)B * C * T * HW;
  const int blocks = (int)min((long long)sm_count() * 16, (total + 255) / 256);
  nhwc_to_ncthw_kernel<<<blocks, 256, 0, stream>>>(x, ldx, out, B, C, T, HW);
  VC_CHECK_CUDA(cudaGetLastError());
  return VC_OK;
}

__global__ void nhwc_to_nchw_h_kernel(const __half* __restrict__ x, int ldx, float* __restrict__ out, int N, int C, long long HW) {
  const long long total = (long long)N * C * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i % HW;
    const long long r = i / HW;
    const int c = (int)(r % C);
    const long long n = r / C;
    out[i] = __half2float(x[(n * HW + p) * ldx + c]);
  }
}
int nhwc_to_nchw_f32_from_f16(const __half* x, int ldx, float* out, int N, int C, long long HW, cudaStream_t stream) {
  VC_REQUIRE(x && out, "nhwc_to_nchw: null pointer");
  const long long total = (long long)N * C * HW;
  const int blocks = (int)min((long long)sm_count() * 16, (total + 255) / 256);
  nhwc_to_nchw_h_kernel<<<blocks, 256, 0, stream>>>(x, ldx, out, N, C, HW);
  VC_CHECK_CUDA(cudaGetLastError());
  return VC_OK;
}

__global__ void cast_kernel(const float* __restrict__ x, __half* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2half_rn(x[i]);
}
int cast_f32_to_f16(const float* x, __half* out, long long n, cudaStream_t stream) {
  VC_REQUIRE(x && out, "cast: null pointer");
  const int blocks = (int)min((long long)sm_count() * 16, (n + 255) / 256);
  cast_kernel<<<blocks, 256, 0, stream>>>(x, out, n);
  VC_CHECK_CUDA(cudaGetLastError());
  return VC_OK;
}

__global__ void add_kernel(const __half2* __restrict__ a, const __half2* __restrict__ b, __half2* __restrict__ out, long long n2) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    const float2 x = __half22float2(a[i]), y = __half22float2(b[i]);
    out[i] = __floats2half2_rn(x.x + y.x, x.y + y.y);
  }
}
// exact-erf GELU, elementwise (Resampler FeedForward, resampler.py:27-34); erf as in the GEGLU epilogue (common.cuh)
__global__ void gelu_kernel(const __half2* __restrict__ x, __half2* __restrict__ out, long long n2) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    const float2 v = __half22float2(x[i]);
    out[i] = __floats2half2_rn(gelu_erf_fast(v.x), gelu_erf_fast(v.y));
  }
}
int gelu_rows_f16(const __half* x, __half* out, long long n, cudaStream_t stream) {
  VC_REQUIRE(x && out && n > 0 && n % 2 == 0, "gelu: bad args");
  const int blocks = (int)min((long long)sm_count() * 16, (n / 2 + 255) / 256);
  gelu_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const __half2*>(x), reinterpret_cast<__half2*>(out), n / 2);
  VC_CHECK_CUDA(cudaGetLastError());
  return VC_OK;
}

int add_rows_f16(const __half* a, const __half* b, __half* out, long long n, cudaStream_t stream) {
  VC_REQUIRE(a && b && out && n % 2 == 0, "add: bad args");
  const int blocks = (int)min((long long)sm_count() * 16, (n / 2 + 255) / 256);
  add_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const __half2*>(a), reinterpret_cast<const __half2*>(b),
                                         reinterpret_cast<__half2*>(out), n / 2);
  VC_CHECK_CUDA(cudaGetLastError());
  return VC_OK;
}

// ------------------------------------------------------------------------------------------------
// time / fps embedding pieces (utils_diffusion.py:8-28, openaimodel3d.py:549-577, :218): fp32, M is 1..B
// ------------------------------------------------------------------------------------------------
__global__ void small_linear_kernel(const float* __restrict__ x, int rows, int K, const float* __restrict__ W,
                                    const float* __restrict__ bias, int N, int act_in, float* __restrict__ out,
                                    const float* __restrict__ add) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows * N) return;
  const int r = warp / N, n = warp % N;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) {
    float xv = x[(long long)r * K + k];
    if (act_in == 1) xv = xv / (1.0f + expf(-xv));
    acc += xv * W[(long long)n * K + k];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    float y = acc + (bias ? bias[n] : 0.f);
    if (add) y += add[(long long)r * N + n];
    out[(long long)r * N + n] = y;
  }
}
int small_linear_f32(const float* x, int rows, int K, const float* W, const float* bias, int N, int act_in, float* out,
                     const float* add, cudaStream_t stream) {
  VC_REQUIRE(x && W && out && rows > 0 && N > 0 && K > 0, "small_linear: bad args");
  const long long warps = (long long)rows * N;
  const int blocks = (int)((warps * 32 + 255) / 256);
  small_linear_kernel<<<blocks, 256, 0, stream>>>(x, rows, K, W, bias, N, act_in, out, add);
  VC_CHECK_CUDA(cudaGetLastError());
  return VC_OK;
}

__global__ void timestep_embedding_kernel(const long long* __restrict__ t, int n, int dim, float* __restrict__ out) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * half) return;
  const int r = i / half, j = i % half;
  // freqs = exp(-ln(10000) * j / half) evaluated in fp32 exactly as torch does
  const float freq = expf(-9.210340371976184f * (float)j / (float)half);
  const float arg = (float)t[r] * freq;
  out[(long long)r * dim + j] = cosf(arg);
  out[(long long)r * dim + half + j] = sinf(arg);
  if ((dim & 1) && j == 0) out[(long long)r * dim + dim - 1] = 0.f;
}
int timestep_embedding_f32(const long long* t, int n, int dim, float* out, cudaStream_t stream) {
  VC_REQUIRE(t && out && n > 0 && dim >= 2, "timestep_embedding: bad args");
  const int total = n * (dim / 2);
  timestep_embedding_kernel<<<(total + 127) / 128, 128, 0, stream>>>(t, n, dim, out);
  VC_CHECK_CUDA(cudaGetLastError());
  return VC_OK;
}

// ------------------------------------------------------------------------------------------------
// Fused DDIM update (ddim.py:228-281 + utils_diffusion.py:147-158), v-parameterisation, batch 1 per call.
// pass 1: double-precision sums for the two unbiased stds; pass 2: elementwise update.
// ------------------------------------------------------------------------------------------------
// CFG combine.  Two-way (ddim.py:226): u + s (c - u).  Three-way (ddim_multiplecond.py:233, vi = the "image, no text"
// branch): u + s_img (vi - u) + s (c - vi), evaluated left to right like the reference expression.
__device__ __forceinline__ float cfg_combine(float c, float u, const float* __restrict__ vi, long long i, float cfg, float cfg_img) {
  if (vi == nullptr) return u + cfg * (c - u);
  const float w = vi[i];
  return (u + cfg_img * (w - u)) + cfg * (c - w);
}
// deterministic: every block leaves its four partial sums in ws[4 + 4 * block] (fixed in-block order); ddim_apply_kernel adds the
// blocks up in index order -- no floating-point atomics, a seeded sampling run is bit-reproducible
static constexpr int DDIM_MAX_BLOCKS = 1024;
__global__ void __launch_bounds__(256) ddim_stats_kernel(const float* __restrict__ vc_, const float* __restrict__ vu, const float* __restrict__ vi, long long n,
                                  float cfg, float cfg_img, double* ws) {
  __shared__ double red[8][4];
  double s1 = 0, q1 = 0, s2 = 0, q2 = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float c = vc_[i], u = vu[i];
    const float m = cfg_combine(c, u, vi, i, cfg, cfg_img);
    s1 += c; q1 += (double)c * c;
    s2 += m; q2 += (double)m * m;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o); q2 += __shfl_xor_sync(0xffffffffu, q2, o);
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[w][0] = s1; red[w][1] = q1; red[w][2] = s2; red[w][3] = q2; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double a = 0;
    for (int i = 0; i < 8; ++i) a += red[i][threadIdx.x];
    ws[4 + 4 * blockIdx.x + threadIdx.x] = a;
  }
}
__global__ void ddim_apply_kernel(const float* __restrict__ x, const float* __restrict__ vc_, const float* __restrict__ vu,
                                  const float* __restrict__ vi, float cfg_img,
                                  const float* __restrict__ noise, float* __restrict__ x_prev, float* __restrict__ pred_x0,
                                  long long n, DdimStepScalars s, const double* ws, int stat_blocks) {
  float factor = 1.f;
  if (s.use_cfg && s.guidance_rescale > 0.f) {
    __shared__ double tot[4];
    if (threadIdx.x < 4) {
      double a = 0;
      for (int b = 0; b < stat_blocks; ++b) a += ws[4 + 4 * b + threadIdx.x];
      tot[threadIdx.x] = a;
    }
    __syncthreads();
    const double dn = (double)n;
    const double var_t = (tot[1] - tot[0] * tot[0] / dn) / (dn - 1.0);
    const double var_c = (tot[3] - tot[2] * tot[2] / dn) / (dn - 1.0);
    factor = (float)sqrt(var_t > 0 ? var_t : 0.0) / (float)sqrt(var_c > 0 ? var_c : 0.0);
  }
  const float rescale = s.prev_scale_t / s.scale_t;
  const float dir_c = sqrtf(1.f - s.a_prev - s.sigma_t * s.sigma_t);
  const float sq_ap = sqrtf(s.a_prev);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float c = vc_[i];
    float m = c;
    if (s.use_cfg) {
      const float u = vu[i];
      m = cfg_combine(c, u, vi, i, s.cfg_scale, cfg_img);
      if (s.guidance_rescale > 0.f) m = s.guidance_rescale * (m * factor) + (1.f - s.guidance_rescale) * m;
    }
    const float xi = x[i];
    const float e_t = s.sqrt_ac_t * m + s.sqrt_1mac_t * xi;
    float p0 = s.sqrt_ac_t * xi - s.sqrt_1mac_t * m;
    p0 *= rescale;
    pred_x0[i] = p0;
    x_prev[i] = sq_ap * p0 + dir_c * e_t + s.sigma_t * noise[i];
  }
}
int ddim_update(const float* x, const float* v_cond, const float* v_uncond, const float* v_uncond_img, float cfg_img, const float* noise,
                float* x_prev, float* pred_x0, long long n, const DdimStepScalars& s, double* ws, cudaStream_t stream) {
  VC_REQUIRE(x && v_cond && noise && x_prev && pred_x0 && ws && n > 1, "ddim_update: bad args");
  VC_REQUIRE(!s.use_cfg || v_uncond, "ddim_update: CFG needs the unconditional output");
  VC_REQUIRE(!v_uncond_img || s.use_cfg, "ddim_update: the image-only branch is only defined with CFG on");
  int blocks = (int)min((long long)sm_count() * 4, (n + 255) / 256);
  if (blocks > DDIM_MAX_BLOCKS) blocks = DDIM_MAX_BLOCKS;
  if (s.use_cfg && s.guidance_rescale > 0.f) {           // ws: 4 * (1 + DDIM_MAX_BLOCKS) doubles
    ddim_stats_kernel<<<blocks, 256, 0, stream>>>(v_cond, v_uncond, v_uncond_img, n, s.cfg_scale, cfg_img, ws);
    VC_CHECK_CUDA(cudaGetLastError());
  }
  ddim_apply_kernel<<<blocks, 256, 0, stream>>>(x, v_cond, v_uncond, v_uncond_img, cfg_img, noise, x_prev, pred_x0, n, s, ws, blocks);
  VC_CHECK_CUDA(cudaGetLastError());
  return VC_OK;
}

// ------------------------------------------------------------------------------------------------
// row softmax (fp32 scores -> fp16 probabilities) for the VAE's single-head d=512 AttnBlock (ae_modules.py:66-68)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ x, long long cols, float scale, __half* __restrict__ out) {
  __shared__ float red[8];
  const float* xr = x + (long long)blockIdx.x * cols;
  __half* orow = out + (long long)blockIdx.x * cols;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  float m = -INFINITY;
  for (long long i = tid; i < cols; i += 256) m = fmaxf(m, xr[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[w] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float l = 0.f;
  for (long long i = tid; i < cols; i += 256) l += __expf((xr[i] - m) * scale);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  if (lane == 0) red[w] = l;
  __syncthreads();
  l = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) l += red[i];
  const float inv = 1.f / l;
  for (long long i = tid; i < cols; i += 256) orow[i] = __float2half_rn(__expf((xr[i] - m) * scale) * inv);
}
int softmax_rows_f32(const float* x, long long rows, long long cols, float scale, __half* out, cudaStream_t stream) {
  VC_REQUIRE(x && out && rows > 0 && cols > 0 && scale > 0.f, "softmax_rows: bad args");
  softmax_rows_kernel<<<(unsigned)rows, 256, 0, stream>>>(x, cols, scale, out);
  VC_CHECK_CUDA(cudaGetLastError());
  return VC_OK;
}

}  // namespace vc
