// HBM-bound kernels of the U-Net step (sm_100a): BatchNorm statistics / apply / backward, 2x2 max-pool with argmax,
// dropout, skip-gradient merge, Adam, weight repacking, dropout-mask generation, z-score normalisation.
// All activations are NHWC; T = __nv_bfloat16 (product path) or float (fp32 check mode).  Every thread moves 8
// consecutive channels (128-bit accesses for bf16); per-channel reductions are thread-private over a grid-stride
// pixel loop, then one shared-memory tree per block, then one partial row per block (deterministic, no atomics).
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int TPB = 256;
#ifndef UB_BN_MINB
#define UB_BN_MINB 3
#endif

template <typename T>
struct V8;
template <>
struct V8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]);
    u.y = pack_bf16x2(f[2], f[3]);
    u.z = pack_bf16x2(f[4], f[5]);
    u.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <>
struct V8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&f)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&f)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
};

__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) { V8<float>::load(p, f); }

// value as it will be read back from storage of type T
template <typename T>
__device__ __forceinline__ float storage_round(float v);
template <>
__device__ __forceinline__ float storage_round<__nv_bfloat16>(float v) { return __bfloat162float(__float2bfloat16(v)); }
template <>
__device__ __forceinline__ float storage_round<float>(float v) { return v; }

// Block-level finish of per-channel partials.  Thread t owns channel group g = t % G (8 channels) and pixel lane
// t / G.  Writes row `blockIdx.x` of partial[rows][NCOMP][C].
template <int NCOMP>
__device__ __forceinline__ void block_channel_reduce(float (&acc)[NCOMP][8], int C, float* __restrict__ partial, size_t row_stride = 0) {
  if (row_stride == 0) row_stride = (size_t)NCOMP * C;
  __shared__ float red[TPB * 8];
  const int G = C >> 3;
  const int PL = TPB / G;
  const int g = threadIdx.x % G, pl = threadIdx.x / G;
#pragma unroll
  for (int comp = 0; comp < NCOMP; ++comp) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) red[pl * C + g * 8 + i] = acc[comp][i];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += TPB) {
      float t = 0.f;
      for (int l = 0; l < PL; ++l) t += red[l * C + c];
      partial[(size_t)blockIdx.x * row_stride + (size_t)comp * C + c] = t;
    }
  }
}

// ------------------------------------------------------------------ BN statistics (standalone)
template <typename T>
__global__ void __launch_bounds__(TPB) bn_stats_kernel(const T* __restrict__ a, float* __restrict__ partial, long long M, int C) {
  const int G = C >> 3, PL = TPB / G;
  const int g = threadIdx.x % G, pl = threadIdx.x / G;
  float acc[2][8] = {};
  for (long long px = (long long)blockIdx.x * PL + pl; px < M; px += (long long)gridDim.x * PL) {
    float f[8];
    V8<T>::load(a + px * C + g * 8, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[0][i] += f[i];
      acc[1][i] += f[i] * f[i];
    }
  }
  block_channel_reduce<2>(acc, C, partial);
}

// Row-parallel column sums: a block owns 32 columns; thread (lane = column, ry = row lane) sums rows ry, ry+8, ... in
// fp64, the 8 row lanes are combined through shared memory in a fixed order (deterministic).
constexpr int RED_ROWLANES = 32;
__device__ __forceinline__ double block_rows_sum(double v, double (*sh)[32]) {
  const int lane = threadIdx.x & 31, ry = threadIdx.x >> 5;
  __syncthreads();
  sh[ry][lane] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int r = 0; r < RED_ROWLANES; ++r) t += sh[r][lane];
  return t;
}

// partial[rows][2][ncols]; channel c gathers columns g*C + c for g < groups.  Biased variance normalises (training),
// the unbiased one feeds the moving average.
__global__ void __launch_bounds__(32 * RED_ROWLANES) bn_finalize_kernel(const float* __restrict__ partial, int rows, int ncols, int groups,
                                                                        double count, float* __restrict__ mean, float* __restrict__ rstd,
                                                                        float* __restrict__ moving_mean, float* __restrict__ moving_var,
                                                                        float momentum, float eps) {
  __shared__ double sh[RED_ROWLANES][32];
  const int C = ncols / groups;
  const int lane = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double s = 0.0, q = 0.0;
  if (c < C) {
    if (rows == UB_STATS_ROWS && groups == 1) {
      // the common case: all loads of this thread are issued before the first add (the kernel is pure L2 latency otherwise)
      constexpr int PER = (UB_STATS_ROWS + RED_ROWLANES - 1) / RED_ROWLANES;
      float vs[PER], vq[PER];
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int r = ry + i * RED_ROWLANES;
        vs[i] = r < UB_STATS_ROWS ? __ldg(partial + ((size_t)r * 2 + 0) * ncols + c) : 0.f;
        vq[i] = r < UB_STATS_ROWS ? __ldg(partial + ((size_t)r * 2 + 1) * ncols + c) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        s += (double)vs[i];
        q += (double)vq[i];
      }
    } else {
#pragma unroll 4
      for (int r = ry; r < rows; r += RED_ROWLANES)
        for (int g = 0; g < groups; ++g) {
          s += (double)partial[((size_t)r * 2 + 0) * ncols + g * C + c];
          q += (double)partial[((size_t)r * 2 + 1) * ncols + g * C + c];
        }
    }
  }
  s = block_rows_sum(s, sh);
  q = block_rows_sum(q, sh);
  if (ry != 0 || c >= C) return;
  const double mu = s / count;
  double var = q / count - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)mu;
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (moving_mean) {
    const double unb = count > 1.0 ? var * (count / (count - 1.0)) : var;
    moving_mean[c] = momentum * moving_mean[c] + (1.f - momentum) * (float)mu;
    moving_var[c] = momentum * moving_var[c] + (1.f - momentum) * (float)unb;
  }
}

// out[c] = scale * sum_rows partial[r * row_stride + c],  c < ncols
__global__ void __launch_bounds__(32 * RED_ROWLANES) reduce_rows_kernel(const float* __restrict__ partial, int rows, int row_stride, int ncols,
                                                                        float* __restrict__ out, float scale) {
  __shared__ double sh[RED_ROWLANES][32];
  const int lane = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double s = 0.0;
  if (c < ncols) {
    if (rows == UB_STATS_ROWS) {
      constexpr int PER = (UB_STATS_ROWS + RED_ROWLANES - 1) / RED_ROWLANES;
      float vs[PER];
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int r = ry + i * RED_ROWLANES;
        vs[i] = r < UB_STATS_ROWS ? __ldg(partial + (size_t)r * row_stride + c) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < PER; ++i) s += (double)vs[i];
    } else {
#pragma unroll 4
      for (int r = ry; r < rows; r += RED_ROWLANES) s += (double)partial[(size_t)r * row_stride + c];
    }
  }
  s = block_rows_sum(s, sh);
  if (ry == 0 && c < ncols) out[c] = (float)(s * (double)scale);
}

// inference fold: scale = gamma / sqrt(var + eps), shift = beta - mean * scale
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                               const float* __restrict__ var, float* __restrict__ scale, float* __restrict__ shift, int n, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float sc = (float)((double)gamma[i] / sqrt((double)var[i] + (double)eps));
    scale[i] = sc;
    shift[i] = beta[i] - mean[i] * sc;
  }
}

__global__ void rsqrt_eps_kernel(const float* __restrict__ v, float* __restrict__ out, int n, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)(1.0 / sqrt((double)v[i] + (double)eps));
}

// ------------------------------------------------------------------ BN apply (+dropout) (+2x2 max-pool)
// y = a*scale + shift with scale = gamma*rstd, shift = beta - mean*scale; optional inverted dropout (x2 on keep).
template <typename T>
__global__ void __launch_bounds__(TPB) bn_apply_kernel(const T* __restrict__ a, T* __restrict__ y, const float* __restrict__ mean,
                                                       const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, const uint8_t* __restrict__ drop_mask,
                                                       long long total8, int C) {
  // the grid stride (gridDim.x * TPB) is a multiple of the channel-group count G (G divides TPB), so a thread keeps
  // its 8 channels for the whole loop and the folded affine lives in registers
  const int G = C >> 3;
  const long long i0 = (long long)blockIdx.x * TPB + threadIdx.x;
  const long long stride = (long long)gridDim.x * TPB;
  const int c0 = (int)(i0 % G) * 8;
  float sc[8], sh[8];
  {
    float mu[8], rs[8], ga[8], be[8];
    load8f(mean + c0, mu); load8f(rstd + c0, rs); load8f(gamma + c0, ga); load8f(beta + c0, be);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      sc[k] = ga[k] * rs[k];
      sh[k] = be[k] - mu[k] * sc[k];
    }
  }
  for (long long i = i0; i < total8; i += 2 * stride) {
    const long long j = i + stride;
    const bool two = j < total8;
    float f[8], g[8];
    uint2 dm0 = make_uint2(0x01010101u, 0x01010101u), dm1 = dm0;
    V8<T>::load(a + i * 8, f);
    if (two) V8<T>::load(a + j * 8, g);
    if (drop_mask) {
      dm0 = __ldg(reinterpret_cast<const uint2*>(drop_mask + i * 8));
      if (two) dm1 = __ldg(reinterpret_cast<const uint2*>(drop_mask + j * 8));
    }
    const uint8_t* d0 = reinterpret_cast<const uint8_t*>(&dm0);
    const uint8_t* d1 = reinterpret_cast<const uint8_t*>(&dm1);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = fmaf(f[k], sc[k], sh[k]);
      if (drop_mask) v = d0[k] ? 2.f * v : 0.f;
      f[k] = v;
    }
    V8<T>::store(y + i * 8, f);
    if (two) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float v = fmaf(g[k], sc[k], sh[k]);
        if (drop_mask) v = d1[k] ? 2.f * v : 0.f;
        g[k] = v;
      }
      V8<T>::store(y + j * 8, g);
    }
  }
}

// One thread = one 2x2 window x 8 channels: writes the 4 normalised pixels, the pooled max and its slot (2*dy+dx,
// first max wins, matching argmax tie-breaking of the oracle).
// STORE_Y = false: only the pooled tensor and the argmax slots are written (the consumers of y fold this BatchNorm, csrc/fold.cu)
template <typename T, bool STORE_Y = true>
__global__ void __launch_bounds__(TPB, 3) bn_apply_pool_kernel(const T* __restrict__ a, T* __restrict__ y, T* __restrict__ pooled,
                                                            uint8_t* __restrict__ idx, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, const uint8_t* __restrict__ drop_mask,
                                                            int N, int H, int W, int C) {
  const int G = C >> 3, Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)N * Ho * Wo * G;
  // grid stride is a multiple of G (G divides TPB): the thread's channel group and folded affine are loop invariant
  const int c0 = (int)(((long long)blockIdx.x * TPB + threadIdx.x) % G) * 8;
  float sc[8], sh[8];
  {
    float mu[8], rs[8], ga[8], be[8];
    load8f(mean + c0, mu); load8f(rstd + c0, rs); load8f(gamma + c0, ga); load8f(beta + c0, be);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      sc[k] = ga[k] * rs[k];
      sh[k] = be[k] - mu[k] * sc[k];
    }
  }
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < total; i += (long long)gridDim.x * TPB) {
    long long t = i / G;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float best[8];
    int slot[8];
    float fa[4][8];
    uint2 dma[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const long long off = (((long long)n * H + 2 * ho + (s >> 1)) * W + 2 * wo + (s & 1)) * C + c0;
      V8<T>::load(a + off, fa[s]);
      dma[s] = make_uint2(0x01010101u, 0x01010101u);
      if (drop_mask) dma[s] = __ldg(reinterpret_cast<const uint2*>(drop_mask + off));
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const long long off = (((long long)n * H + 2 * ho + (s >> 1)) * W + 2 * wo + (s & 1)) * C + c0;
      float (&f)[8] = fa[s];
      const uint8_t* dmb = reinterpret_cast<const uint8_t*>(&dma[s]);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float v = fmaf(f[k], sc[k], sh[k]);
        if (drop_mask) v = dmb[k] ? 2.f * v : 0.f;
        f[k] = v;
      }
      if (STORE_Y) V8<T>::store(y + off, f);
      // compare what the consumer will read (storage precision), so the saved slot agrees with a re-computed pool
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float r = storage_round<T>(f[k]);
        if (s == 0 || r > best[k]) {
          best[k] = r;
          slot[k] = s;
        }
      }
    }
    const long long po = (((long long)n * Ho + ho) * Wo + wo) * C + c0;
    V8<T>::store(pooled + po, best);
    uint2 pk;
    pk.x = slot[0] | (slot[1] << 8) | (slot[2] << 16) | (slot[3] << 24);
    pk.y = slot[4] | (slot[5] << 8) | (slot[6] << 16) | (slot[7] << 24);
    *reinterpret_cast<uint2*>(idx + po) = pk;
  }
}

// dy[n,h,w,c] = (slot matches ? dpool : 0) + dskip   (skip fan-out sum), optional dropout backward.
// RED: dy is dL/dy of a BatchNorm'd tensor whose saved activation is `a`: the BatchNorm-backward sums of that layer
// (red[row][0][c] = sum dy, red[row][1][c] = rstd_c * sum dy * (a - mean_c), of the STORED dy) are accumulated here, which
// replaces the separate bn_bwd_reduce pass over dy and a.
template <typename T, bool RED>
__global__ void __launch_bounds__(TPB) pool_bwd_add_kernel(const T* __restrict__ dpool, const uint8_t* __restrict__ idx,
                                                           const T* __restrict__ dskip, const uint8_t* __restrict__ drop_mask,
                                                           T* __restrict__ dy, int N, int H, int W, int C, const T* __restrict__ a,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           float* __restrict__ red) {
  const int G = C >> 3, Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)N * Ho * Wo * G;
  float acc[2][8] = {};
  float mu[8] = {};
  if (RED) load8f(mean + (threadIdx.x % G) * 8, mu);      // gridDim.x * TPB is a multiple of G: a thread keeps its channel group
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < total; i += (long long)gridDim.x * TPB) {
    const int g = (int)(i % G);
    long long t = i / G;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    const int c0 = g * 8;
    const long long po = (((long long)n * Ho + ho) * Wo + wo) * C + c0;
    float dp[8];
    V8<T>::load(dpool + po, dp);
    const uint2 pk = __ldg(reinterpret_cast<const uint2*>(idx + po));
    const uint8_t* sl = reinterpret_cast<const uint8_t*>(&pk);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const long long off = (((long long)n * H + 2 * ho + (s >> 1)) * W + 2 * wo + (s & 1)) * C + c0;
      float f[8], av[8];
      if (dskip) V8<T>::load(dskip + off, f);
      else {
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = 0.f;
      }
      if (RED) V8<T>::load(a + off, av);
      uint2 dm = make_uint2(0x01010101u, 0x01010101u);
      if (drop_mask) dm = __ldg(reinterpret_cast<const uint2*>(drop_mask + off));
      const uint8_t* dmb = reinterpret_cast<const uint8_t*>(&dm);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float v = f[k] + (sl[k] == s ? dp[k] : 0.f);
        if (drop_mask) v = dmb[k] ? 2.f * v : 0.f;
        f[k] = v;
        if (RED) {
          const float vs = storage_round<T>(v);
          acc[0][k] += vs;
          acc[1][k] = fmaf(vs, av[k] - mu[k], acc[1][k]);
        }
      }
      V8<T>::store(dy + off, f);
    }
  }
  if (RED) {
    float rs[8];
    load8f(rstd + (threadIdx.x % G) * 8, rs);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[1][k] *= rs[k];
    block_channel_reduce<2>(acc, C, red);
  }
}

// out = in * (mask ? 2 : 0)   (dropout backward where no pool follows: the bottleneck)
template <typename T>
__global__ void __launch_bounds__(TPB) dropout_scale_kernel(const T* __restrict__ in, const uint8_t* __restrict__ mask, T* __restrict__ out,
                                                            long long total8) {
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < total8; i += (long long)gridDim.x * TPB) {
    float f[8];
    V8<T>::load(in + i * 8, f);
    const uint2 dm = __ldg(reinterpret_cast<const uint2*>(mask + i * 8));
    const uint8_t* dmb = reinterpret_cast<const uint8_t*>(&dm);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = dmb[k] ? 2.f * f[k] : 0.f;
    V8<T>::store(out + i * 8, f);
  }
}

// ------------------------------------------------------------------ BN backward
// pass 1: partial[row][0][c] = sum dy, partial[row][1][c] = sum dy * xhat
template <typename T>
__global__ void __launch_bounds__(TPB, UB_BN_MINB) bn_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ a, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, float* __restrict__ partial, long long M, int C) {
  const int G = C >> 3, PL = TPB / G;
  const int g = threadIdx.x % G, pl = threadIdx.x / G;
  float mu[8];
  load8f(mean + g * 8, mu);
  float acc[2][8] = {};
  const long long stride = (long long)gridDim.x * PL;
  for (long long px = (long long)blockIdx.x * PL + pl; px < M; px += 2 * stride) {
    const long long px2 = px + stride;
    const bool two = px2 < M;
    float d[8], f[8], d2[8] = {}, f2[8] = {};
    V8<T>::load(dy + px * C + g * 8, d);
    V8<T>::load(a + px * C + g * 8, f);
    if (two) {
      V8<T>::load(dy + px2 * C + g * 8, d2);
      V8<T>::load(a + px2 * C + g * 8, f2);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[0][i] += d[i];
      acc[1][i] = fmaf(d[i], f[i] - mu[i], acc[1][i]);
    }
    if (two) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[0][i] += d2[i];
        acc[1][i] = fmaf(d2[i], f2[i] - mu[i], acc[1][i]);
      }
    }
  }
  {
    float rs[8];
    load8f(rstd + g * 8, rs);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[1][i] *= rs[i];      // sum dy * xhat = rstd * sum dy * (a - mean)
  }
  block_channel_reduce<2>(acc, C, partial);
}

// pass 2: dz = gamma*rstd*(dy - dbeta/M - xhat*dgamma/M) * [a > 0 if relu];  partial[row][0][c] = sum dz (bias gradient)
template <typename T>
__global__ void __launch_bounds__(TPB, UB_BN_MINB) bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ a, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ dbeta, const float* __restrict__ dgamma,
                                                           T* __restrict__ dz, float* __restrict__ partial, long long M, int C, int relu) {
  const int G = C >> 3, PL = TPB / G;
  const int g = threadIdx.x % G, pl = threadIdx.x / G;
  // dz = gamma*rstd*(dy - dbeta/M - xhat*dgamma/M) = cA*dy + cB*a + cC  with xhat = (a - mean)*rstd
  float cA[8], cB[8], cC[8];
  {
    float mu[8], rs[8], ga[8], db[8], dg[8];
    load8f(mean + g * 8, mu); load8f(rstd + g * 8, rs); load8f(gamma + g * 8, ga);
    load8f(dbeta + g * 8, db); load8f(dgamma + g * 8, dg);
    const float invM = 1.f / (float)M;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      cA[i] = ga[i] * rs[i];
      cB[i] = -cA[i] * rs[i] * dg[i] * invM;
      cC[i] = -cA[i] * db[i] * invM - cB[i] * mu[i];
    }
  }
  float acc[1][8] = {};
  const long long stride = (long long)gridDim.x * PL;
  for (long long px = (long long)blockIdx.x * PL + pl; px < M; px += 2 * stride) {
    const long long px2 = px + stride;
    const bool two = px2 < M;
    float d[8], f[8], d2[8] = {}, f2[8] = {};
    V8<T>::load(dy + px * C + g * 8, d);
    V8<T>::load(a + px * C + g * 8, f);
    if (two) {
      V8<T>::load(dy + px2 * C + g * 8, d2);
      V8<T>::load(a + px2 * C + g * 8, f2);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = fmaf(cA[i], d[i], fmaf(cB[i], f[i], cC[i]));
      if (relu && !(f[i] > 0.f)) v = 0.f;
      d[i] = v;
      acc[0][i] += v;   // bias gradient from the fp32 value (the reference is fp32; a deconv bias gradient is analytically 0)
    }
    V8<T>::store(dz + px * C + g * 8, d);
    if (two) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v = fmaf(cA[i], d2[i], fmaf(cB[i], f2[i], cC[i]));
        if (relu && !(f2[i] > 0.f)) v = 0.f;
        d2[i] = v;
        acc[0][i] += v;
      }
      V8<T>::store(dz + px2 * C + g * 8, d2);
    }
  }
  block_channel_reduce<1>(acc, C, partial);
}

// ------------------------------------------------------------------ Adam (Keras formula)
__global__ void __launch_bounds__(TPB) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                   __nv_bfloat16* __restrict__ shadow, long long n, float lr_t, const float* __restrict__ lr_t_dev,
                                                   float b1, float b2, float eps, float gscale) {
  if (lr_t_dev) lr_t = *lr_t_dev;          // step-dependent scalar kept on the device so that a captured CUDA graph can be replayed
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < n4; i += (long long)gridDim.x * TPB) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = ga[k] * gscale;
      ma[k] = b1 * ma[k] + (1.f - b1) * gk;
      va[k] = b2 * va[k] + (1.f - b2) * gk * gk;
      pa[k] -= lr_t * ma[k] / (sqrtf(va[k]) + eps);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (shadow) {
      uint2 s;
      s.x = pack_bf16x2(pp.x, pp.y);
      s.y = pack_bf16x2(pp.z, pp.w);
      reinterpret_cast<uint2*>(shadow)[i] = s;
    }
  }
  // tail (n not a multiple of 4)
  const long long i = n4 * 4 + (long long)blockIdx.x * TPB + threadIdx.x;
  if (blockIdx.x == 0 && i < n) {
    const float gk = g[i] * gscale;
    m[i] = b1 * m[i] + (1.f - b1) * gk;
    v[i] = b2 * v[i] + (1.f - b2) * gk * gk;
    p[i] -= lr_t * m[i] / (sqrtf(v[i]) + eps);
    if (shadow) shadow[i] = __float2bfloat16(p[i]);
  }
}

// dst[c][t'][r] = src(r, t, c), t' = flip ? T-1-t : t;  dst (bf16 or fp32) is [C][T][R].
// src_layout 0: src fp32 [R][T][C] (conv kernels [Cout][tap][Cin]); 1: src fp32 [T][R][C] (deconv kernels [(ab)][Cout][Cin])
template <typename TO>
__global__ void transpose_pack_kernel(const float* __restrict__ src, TO* __restrict__ dst, int R, int T, int C, int flip, int src_layout) {
  __shared__ float tile[32][33];
  const int t = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    const size_t si = src_layout ? ((size_t)t * R + r) * C + c : ((size_t)r * T + t) * C + c;
    tile[j][threadIdx.x] = (r < R && c < C) ? src[si] : 0.f;
  }
  __syncthreads();
  const int tt = flip ? T - 1 - t : t;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < R && c < C) dst[((size_t)c * T + tt) * R + r] = (TO)tile[threadIdx.x][j];
  }
}

// All dgrad packs of the model in ONE launch: job j transposes src_j (fp32, layout as above) into dst_j; blockIdx.x walks the
// concatenated 32x32 tile lists (tile_begin[j] .. tile_begin[j+1]).
struct PackJob {
  const float* src;
  void* dst;
  int R, T, C, flip, src_layout, tiles_c, tiles_r;
  int tile_begin;
};
template <typename TO>
__global__ void transpose_pack_multi_kernel(const PackJob* __restrict__ jobs, int njobs) {
  __shared__ float tile[32][33];
  int lo = 0, hi = njobs - 1;
  const int b = blockIdx.x;
  while (lo < hi) {                      // last job whose tile_begin <= b
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].tile_begin <= b) lo = mid; else hi = mid - 1;
  }
  const PackJob J = jobs[lo];
  int t = b - J.tile_begin;
  const int tc = t % J.tiles_c; t /= J.tiles_c;
  const int tr = t % J.tiles_r;
  const int tap = t / J.tiles_r;
  const int r0 = tr * 32, c0 = tc * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    const size_t si = J.src_layout ? ((size_t)tap * J.R + r) * J.C + c : ((size_t)r * J.T + tap) * J.C + c;
    tile[j][threadIdx.x] = (r < J.R && c < J.C) ? J.src[si] : 0.f;
  }
  __syncthreads();
  const int tt = J.flip ? J.T - 1 - tap : tap;
  TO* dst = reinterpret_cast<TO*>(J.dst);
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < J.R && c < J.C) dst[((size_t)c * J.T + tt) * J.R + r] = (TO)tile[threadIdx.x][j];
  }
}

__global__ void __launch_bounds__(TPB) cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < n; i += (long long)gridDim.x * TPB) dst[i] = __float2bfloat16(src[i]);
}

// ------------------------------------------------------------------ dropout mask (Philox4x32-10, keep prob 0.5)
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
// one thread -> 16 mask bytes (128 random bits would give 128 masks; we spend 8 bits per element for simplicity)
__global__ void __launch_bounds__(TPB) dropout_mask_kernel(uint8_t* __restrict__ mask, long long n16, unsigned long long seed,
                                                           unsigned long long offset) {
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < n16; i += (long long)gridDim.x * TPB) {
    const unsigned long long ctr = offset + (unsigned long long)i;
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint4 out;
    uint32_t* o = &out.x;
#pragma unroll
    for (int w = 0; w < 4; ++w) o[w] = (c[w] & 0x01010101u);   // bit 0 of each byte: Bernoulli(0.5)
    reinterpret_cast<uint4*>(mask)[i] = out;
  }
}

// ------------------------------------------------------------------ z-score (UNet/imagereader.py:33-66)
// stats: per (image, channel) plane sum and sum of squares in fp64; input u8 / u16 / f32 planes (NCHW).
template <typename TI>
__global__ void __launch_bounds__(TPB) zscore_stats_kernel(const TI* __restrict__ x, double* __restrict__ sums, long long plane) {
  const TI* xp = x + (long long)blockIdx.y * plane;
  double s = 0.0, q = 0.0;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < plane; i += (long long)gridDim.x * TPB) {
    const double v = (double)(float)xp[i];
    s += v;
    q += v * v;
  }
  __shared__ double sh[2][TPB / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = s;
    sh[1][threadIdx.x >> 5] = q;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    for (int w = 0; w < TPB / 32; ++w) {
      ts += sh[0][w];
      tq += sh[1][w];
    }
    // [plane][block][2] partials; summed in fixed order by the apply kernel
    sums[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 + 0] = ts;
    sums[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 + 1] = tq;
  }
}
template <typename TI>
__global__ void __launch_bounds__(TPB) zscore_apply_kernel(const TI* __restrict__ x, float* __restrict__ y, const double* __restrict__ sums,
                                                           int nblk, long long plane) {
  __shared__ float s_mu, s_inv;
  __shared__ double sh[2][TPB / 32];
  {
    // every block reduces the plane's partials the same way (fixed shuffle tree, then warp totals in order): one load per
    // thread instead of a serial loop on thread 0, which dominated this kernel for tile-sized planes
    double ts = 0.0, tq = 0.0;
    for (int b = threadIdx.x; b < nblk; b += TPB) {
      ts += sums[((size_t)blockIdx.y * nblk + b) * 2 + 0];
      tq += sums[((size_t)blockIdx.y * nblk + b) * 2 + 1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ts += __shfl_xor_sync(0xffffffffu, ts, o);
      tq += __shfl_xor_sync(0xffffffffu, tq, o);
    }
    if ((threadIdx.x & 31) == 0) {
      sh[0][threadIdx.x >> 5] = ts;
      sh[1][threadIdx.x >> 5] = tq;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    for (int w = 0; w < TPB / 32; ++w) {
      ts += sh[0][w];
      tq += sh[1][w];
    }
    const double mu = ts / (double)plane;
    double var = tq / (double)plane - mu * mu;
    if (var < 0.0) var = 0.0;
    const float sd = (float)sqrt(var);
    s_mu = (float)mu;
    s_inv = (sd <= 1.0f) ? 1.0f : sd;          // std <= 1 -> subtract the mean only (divide by 1)
  }
  __syncthreads();
  const TI* xp = x + (long long)blockIdx.y * plane;
  float* yp = y + (long long)blockIdx.y * plane;
  const float mu = s_mu, inv = s_inv;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < plane; i += (long long)gridDim.x * TPB)
    yp[i] = ((float)xp[i] - mu) / inv;
}

// ---- z-score in two halves (image split across ranks: each rank sums the rows it owns, the sums are all-reduced, then
// every rank normalises its band with the GLOBAL statistics; unetb200.inference.segment_sharded) -----------------------------
template <typename TI>
__global__ void __launch_bounds__(TPB) zscore_part_stats_kernel(const TI* __restrict__ x, double* __restrict__ partial, long long plane,
                                                                long long plane_stride) {
  const TI* xp = x + (long long)blockIdx.y * plane_stride;
  double s = 0.0, q = 0.0;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < plane; i += (long long)gridDim.x * TPB) {
    const double v = (double)(float)xp[i];
    s += v;
    q += v * v;
  }
  __shared__ double sh[2][TPB / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = s;
    sh[1][threadIdx.x >> 5] = q;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    for (int w = 0; w < TPB / 32; ++w) {
      ts += sh[0][w];
      tq += sh[1][w];
    }
    partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 + 0] = ts;
    partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 + 1] = tq;
  }
}

// sums[plane][2] = fixed-order sum of the plane's block partials (one block per plane)
__global__ void __launch_bounds__(TPB) zscore_fold_partials_kernel(const double* __restrict__ partial, double* __restrict__ sums, int nblk) {
  __shared__ double sh[2][TPB / 32];
  double ts = 0.0, tq = 0.0;
  for (int b = threadIdx.x; b < nblk; b += TPB) {
    ts += partial[((size_t)blockIdx.x * nblk + b) * 2 + 0];
    tq += partial[((size_t)blockIdx.x * nblk + b) * 2 + 1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ts += __shfl_xor_sync(0xffffffffu, ts, o);
    tq += __shfl_xor_sync(0xffffffffu, tq, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = ts;
    sh[1][threadIdx.x >> 5] = tq;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < TPB / 32; ++w) {
      a += sh[0][w];
      b += sh[1][w];
    }
    sums[blockIdx.x * 2 + 0] = a;
    sums[blockIdx.x * 2 + 1] = b;
  }
}

// y = (x - mean) / (std <= 1 ? 1 : std) with mean / std from sums[plane] = (sum, sum of squares) over `count` samples
template <typename TI>
__global__ void __launch_bounds__(TPB) zscore_apply_sums_kernel(const TI* __restrict__ x, float* __restrict__ y, const double* __restrict__ sums,
                                                                double count, long long plane) {
  const double mu_d = sums[blockIdx.y * 2 + 0] / count;
  double var = sums[blockIdx.y * 2 + 1] / count - mu_d * mu_d;
  if (var < 0.0) var = 0.0;
  const float sd = (float)sqrt(var);
  const float mu = (float)mu_d, inv = (sd <= 1.0f) ? 1.0f : sd;
  const TI* xp = x + (long long)blockIdx.y * plane;
  float* yp = y + (long long)blockIdx.y * plane;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < plane; i += (long long)gridDim.x * TPB)
    yp[i] = ((float)xp[i] - mu) / inv;
}

inline int grid_for(long long work_items, int per_block, int cap) {
  long long g = (work_items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

bool channels_ok(int C) { return C >= 64 && C <= 2048 && (C & (C - 1)) == 0; }

}  // namespace

#define UB_DISPATCH_T(dtype, ...)                                   \
  do {                                                              \
    if ((dtype) == UB_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else if ((dtype) == UB_F32) { using T = float; __VA_ARGS__; }   \
    else { ub_set_error("bad dtype %d", (int)(dtype)); return UB_ERR_INVALID_ARG; } \
  } while (0)

extern "C" {

int ub_bn_stats(const void* a, float* partial, long long M, int C, int dtype, cudaStream_t stream) {
  UB_CHECK_ARG(a && partial && M > 0, "bn_stats: bad args");
  UB_CHECK_SHAPE(channels_ok(C), "bn_stats: C=%d must be a power of two in [64,2048]", C);
  UB_CUDA(cudaMemsetAsync(partial, 0, sizeof(float) * UB_STATS_ROWS * 2 * C, stream));
  const int PL = TPB / (C >> 3);
  const int grid = grid_for(M, PL * 4, UB_STATS_ROWS);
  UB_DISPATCH_T(dtype, (bn_stats_kernel<T><<<grid, TPB, 0, stream>>>((const T*)a, partial, M, C)));
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_bn_finalize(const float* partial, int ncols, int groups, long long count, float* mean, float* rstd, float* moving_mean,
                   float* moving_var, float momentum, float eps, cudaStream_t stream) {
  UB_CHECK_ARG(partial && mean && rstd && groups > 0 && ncols % groups == 0 && count > 0, "bn_finalize: bad args");
  const int C = ncols / groups;
  bn_finalize_kernel<<<(C + 31) / 32, 32 * RED_ROWLANES, 0, stream>>>(partial, UB_STATS_ROWS, ncols, groups, (double)count, mean, rstd,
                                                                     moving_mean, moving_var, momentum, eps);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_reduce_rows(const float* partial, int rows, int row_stride, int ncols, float* out, float scale, cudaStream_t stream) {
  UB_CHECK_ARG(partial && out && rows > 0 && ncols > 0 && row_stride >= ncols, "reduce_rows: bad args");
  reduce_rows_kernel<<<(ncols + 31) / 32, 32 * RED_ROWLANES, 0, stream>>>(partial, rows, row_stride, ncols, out, scale);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_bn_inference_rstd(const float* moving_var, float* rstd, int n, float eps, cudaStream_t stream) {
  UB_CHECK_ARG(moving_var && rstd && n > 0, "bn_inference_rstd: bad args");
  rsqrt_eps_kernel<<<(n + 255) / 256, 256, 0, stream>>>(moving_var, rstd, n, eps);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_bn_fold(const float* gamma, const float* beta, const float* moving_mean, const float* moving_var, float* scale, float* shift, int n,
               float eps, cudaStream_t stream) {
  UB_CHECK_ARG(gamma && beta && moving_mean && moving_var && scale && shift && n > 0, "bn_fold: bad args");
  bn_fold_kernel<<<(n + 255) / 256, 256, 0, stream>>>(gamma, beta, moving_mean, moving_var, scale, shift, n, eps);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_bn_apply(const void* a, void* y, const float* mean, const float* rstd, const float* gamma, const float* beta,
                const unsigned char* drop_mask, long long M, int C, int dtype, cudaStream_t stream) {
  UB_CHECK_ARG(a && y && mean && rstd && gamma && beta && M > 0, "bn_apply: bad args");
  UB_CHECK_SHAPE(channels_ok(C), "bn_apply: C=%d must be a power of two in [64,2048]", C);
  const long long total8 = M * (C / 8);
  const int grid = grid_for(total8, TPB, ub_num_sms() * 16);
  UB_DISPATCH_T(dtype, (bn_apply_kernel<T><<<grid, TPB, 0, stream>>>((const T*)a, (T*)y, mean, rstd, gamma, beta, drop_mask, total8, C)));
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_bn_apply_pool(const void* a, void* y, void* pooled, unsigned char* idx, const float* mean, const float* rstd,
                     const float* gamma, const float* beta, const unsigned char* drop_mask, int N, int H, int W, int C, int dtype,
                     cudaStream_t stream) {
  UB_CHECK_ARG(a && y && pooled && idx && mean && rstd && gamma && beta, "bn_apply_pool: bad args");
  UB_CHECK_SHAPE(channels_ok(C) && H % 2 == 0 && W % 2 == 0, "bn_apply_pool: C=%d must be a power of two in [64,2048], H/W even", C);
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
  const int grid = grid_for(total, TPB, ub_num_sms() * 12);
  UB_DISPATCH_T(dtype, (bn_apply_pool_kernel<T><<<grid, TPB, 0, stream>>>((const T*)a, (T*)y, (T*)pooled, idx, mean, rstd, gamma, beta,
                                                                         drop_mask, N, H, W, C)));
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_bn_pool(const void* a, void* pooled, unsigned char* idx, const float* mean, const float* rstd, const float* gamma,
               const float* beta, const unsigned char* drop_mask, int N, int H, int W, int C, int dtype, cudaStream_t stream) {
  UB_CHECK_ARG(a && pooled && idx && mean && rstd && gamma && beta, "bn_pool: bad args");
  UB_CHECK_SHAPE(channels_ok(C) && H % 2 == 0 && W % 2 == 0, "bn_pool: C=%d must be a power of two in [64,2048], H/W even", C);
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
  const int grid = grid_for(total, TPB, ub_num_sms() * 12);
  UB_DISPATCH_T(dtype, (bn_apply_pool_kernel<T, false><<<grid, TPB, 0, stream>>>((const T*)a, nullptr, (T*)pooled, idx, mean, rstd, gamma, beta,
                                                                                drop_mask, N, H, W, C)));
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_maxpool2x2_bwd_add(const void* dpool, const unsigned char* idx, const void* dskip, const unsigned char* drop_mask, void* dy,
                          int N, int H, int W, int C, int dtype, cudaStream_t stream) {
  UB_CHECK_ARG(dpool && idx && dy, "maxpool2x2_bwd_add: bad args");
  UB_CHECK_SHAPE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool2x2_bwd_add: C %% 8, even H/W");
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
  const int grid = grid_for(total, TPB, ub_num_sms() * 16);
  UB_DISPATCH_T(dtype, (pool_bwd_add_kernel<T, false><<<grid, TPB, 0, stream>>>((const T*)dpool, idx, (const T*)dskip, drop_mask, (T*)dy, N, H, W,
                                                                               C, nullptr, nullptr, nullptr, nullptr)));
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_maxpool2x2_bwd_add_bnred(const void* dpool, const unsigned char* idx, const void* dskip, const unsigned char* drop_mask, void* dy,
                                int N, int H, int W, int C, const void* a, const float* mean, const float* rstd, float* red_partial,
                                int dtype, cudaStream_t stream) {
  UB_CHECK_ARG(dpool && idx && dy && a && mean && rstd && red_partial, "maxpool2x2_bwd_add_bnred: bad args");
  UB_CHECK_SHAPE(channels_ok(C) && H % 2 == 0 && W % 2 == 0, "maxpool2x2_bwd_add_bnred: C=%d must be a power of two in [64,2048], even H/W", C);
  UB_CUDA(cudaMemsetAsync(red_partial, 0, sizeof(float) * UB_STATS_ROWS * 2 * C, stream));
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
  const int cap = ub_num_sms() * 3 < UB_STATS_ROWS ? ub_num_sms() * 3 : UB_STATS_ROWS;      // one wave of resident blocks, one partial row each
  const int grid = grid_for(total, TPB, cap);
  UB_DISPATCH_T(dtype, (pool_bwd_add_kernel<T, true><<<grid, TPB, 0, stream>>>((const T*)dpool, idx, (const T*)dskip, drop_mask, (T*)dy, N, H, W,
                                                                              C, (const T*)a, mean, rstd, red_partial)));
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_dropout_bwd(const void* in, const unsigned char* mask, void* out, long long n, int dtype, cudaStream_t stream) {
  UB_CHECK_ARG(in && mask && out && n % 8 == 0, "dropout_bwd: bad args");
  const int grid = grid_for(n / 8, TPB, ub_num_sms() * 16);
  UB_DISPATCH_T(dtype, (dropout_scale_kernel<T><<<grid, TPB, 0, stream>>>((const T*)in, mask, (T*)out, n / 8)));
  UB_LAUNCH_CHECK();
  return UB_OK;
}

// resident blocks per SM for the two BatchNorm-backward passes (3 fills the SM when they run alone; 2 leaves room for a
// co-resident weight-gradient CTA when the step overlaps them, UB_BN_BWD_CTAS=2)
static int bn_bwd_ctas_per_sm() {
  static int v = 0;
  if (!v) {
    const char* e = getenv("UB_BN_BWD_CTAS");
    v = (e && e[0] >= '1' && e[0] <= '4') ? e[0] - '0' : 3;
  }
  return v;
}

int ub_bn_bwd_reduce(const void* dy, const void* a, const float* mean, const float* rstd, float* partial, long long M, int C, int dtype,
                     cudaStream_t stream) {
  UB_CHECK_ARG(dy && a && mean && rstd && partial && M > 0, "bn_bwd_reduce: bad args");
  UB_CHECK_SHAPE(channels_ok(C), "bn_bwd_reduce: C=%d must be a power of two in [64,2048]", C);
  UB_CUDA(cudaMemsetAsync(partial, 0, sizeof(float) * UB_STATS_ROWS * 2 * C, stream));
  const int PL = TPB / (C >> 3);
  const int grid = grid_for(M, PL * 4, ub_num_sms() * bn_bwd_ctas_per_sm());      // one wave of resident blocks (<= UB_STATS_ROWS)
  UB_DISPATCH_T(dtype, (bn_bwd_reduce_kernel<T><<<grid, TPB, 0, stream>>>((const T*)dy, (const T*)a, mean, rstd, partial, M, C)));
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_bn_bwd_apply(const void* dy, const void* a, const float* mean, const float* rstd, const float* gamma, const float* dbeta,
                    const float* dgamma, void* dz, float* partial, long long M, int C, int relu, int dtype, cudaStream_t stream) {
  UB_CHECK_ARG(dy && a && mean && rstd && gamma && dbeta && dgamma && dz && partial && M > 0, "bn_bwd_apply: bad args");
  UB_CHECK_SHAPE(channels_ok(C), "bn_bwd_apply: C=%d must be a power of two in [64,2048]", C);
  UB_CUDA(cudaMemsetAsync(partial, 0, sizeof(float) * UB_STATS_ROWS * C, stream));
  const int PL = TPB / (C >> 3);
  const int grid = grid_for(M, PL * 4, ub_num_sms() * bn_bwd_ctas_per_sm());
  UB_DISPATCH_T(dtype, (bn_bwd_apply_kernel<T><<<grid, TPB, 0, stream>>>((const T*)dy, (const T*)a, mean, rstd, gamma, dbeta, dgamma, (T*)dz,
                                                                        partial, M, C, relu)));
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_adam(float* param, const float* grad, float* m, float* v, void* bf16_shadow, long long n, float lr_t, float beta1, float beta2,
            float eps, float grad_scale, cudaStream_t stream) {
  UB_CHECK_ARG(param && grad && m && v && n > 0, "adam: bad args");
  const int grid = grid_for(n / 4 + 1, TPB, ub_num_sms() * 8);
  adam_kernel<<<grid, TPB, 0, stream>>>(param, grad, m, v, (__nv_bfloat16*)bf16_shadow, n, lr_t, nullptr, beta1, beta2, eps, grad_scale);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_adam_dev(float* param, const float* grad, float* m, float* v, void* bf16_shadow, long long n, const float* lr_t_dev, float beta1,
                float beta2, float eps, float grad_scale, cudaStream_t stream) {
  UB_CHECK_ARG(param && grad && m && v && lr_t_dev && n > 0, "adam_dev: bad args");
  const int grid = grid_for(n / 4 + 1, TPB, ub_num_sms() * 8);
  adam_kernel<<<grid, TPB, 0, stream>>>(param, grad, m, v, (__nv_bfloat16*)bf16_shadow, n, 0.f, lr_t_dev, beta1, beta2, eps, grad_scale);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_transpose_pack(const float* src, void* dst, int R, int T, int C, int flip, int src_layout, int dst_dtype, cudaStream_t stream) {
  UB_CHECK_ARG(src && dst && R > 0 && T > 0 && C > 0, "transpose_pack: bad args");
  dim3 grid((C + 31) / 32, (R + 31) / 32, T), block(32, 8);
  if (dst_dtype == UB_BF16) transpose_pack_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>(src, (__nv_bfloat16*)dst, R, T, C, flip, src_layout);
  else if (dst_dtype == UB_F32) transpose_pack_kernel<float><<<grid, block, 0, stream>>>(src, (float*)dst, R, T, C, flip, src_layout);
  else { ub_set_error("transpose_pack: bad dtype"); return UB_ERR_INVALID_ARG; }
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_transpose_pack_multi(const void* jobs_dev, int njobs, int total_tiles, int dst_dtype, cudaStream_t stream) {
  UB_CHECK_ARG(jobs_dev && njobs > 0 && total_tiles > 0, "transpose_pack_multi: bad args");
  dim3 block(32, 8);
  if (dst_dtype == UB_BF16) transpose_pack_multi_kernel<__nv_bfloat16><<<total_tiles, block, 0, stream>>>((const PackJob*)jobs_dev, njobs);
  else if (dst_dtype == UB_F32) transpose_pack_multi_kernel<float><<<total_tiles, block, 0, stream>>>((const PackJob*)jobs_dev, njobs);
  else { ub_set_error("transpose_pack_multi: bad dtype"); return UB_ERR_INVALID_ARG; }
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_cast_bf16(const float* src, void* dst, long long n, cudaStream_t stream) {
  UB_CHECK_ARG(src && dst && n > 0, "cast_bf16: bad args");
  cast_bf16_kernel<<<grid_for(n, TPB, ub_num_sms() * 8), TPB, 0, stream>>>(src, (__nv_bfloat16*)dst, n);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_dropout_mask(unsigned char* mask, long long n, unsigned long long seed, unsigned long long offset, cudaStream_t stream) {
  UB_CHECK_ARG(mask && n > 0 && n % 16 == 0, "dropout_mask: n must be a multiple of 16");
  dropout_mask_kernel<<<grid_for(n / 16, TPB, ub_num_sms() * 8), TPB, 0, stream>>>(mask, n / 16, seed, offset);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

// src planes [planes][plane] of src_dtype (0 = u8, 1 = u16, 2 = f32) -> dst fp32, each plane z-scored on its own.
// scratch: at least planes * UB_ZSCORE_BLOCKS * 2 doubles.
int ub_zscore(const void* src, int src_dtype, float* dst, double* scratch, int planes, long long plane, cudaStream_t stream) {
  UB_CHECK_ARG(src && dst && scratch && planes > 0 && plane > 0, "zscore: bad args");
  const int nblk = grid_for(plane, TPB * 8, UB_ZSCORE_BLOCKS);
  dim3 grid(nblk, planes);
  dim3 agrid(grid_for(plane, TPB * 4, 4096), planes);
  if (src_dtype == 0) {
    zscore_stats_kernel<uint8_t><<<grid, TPB, 0, stream>>>((const uint8_t*)src, scratch, plane);
    zscore_apply_kernel<uint8_t><<<agrid, TPB, 0, stream>>>((const uint8_t*)src, dst, scratch, nblk, plane);
  } else if (src_dtype == 1) {
    zscore_stats_kernel<uint16_t><<<grid, TPB, 0, stream>>>((const uint16_t*)src, scratch, plane);
    zscore_apply_kernel<uint16_t><<<agrid, TPB, 0, stream>>>((const uint16_t*)src, dst, scratch, nblk, plane);
  } else if (src_dtype == 2) {
    zscore_stats_kernel<float><<<grid, TPB, 0, stream>>>((const float*)src, scratch, plane);
    zscore_apply_kernel<float><<<agrid, TPB, 0, stream>>>((const float*)src, dst, scratch, nblk, plane);
  } else {
    ub_set_error("zscore: bad src dtype %d", src_dtype);
    return UB_ERR_INVALID_ARG;
  }
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_zscore_sums(const void* src, int src_dtype, double* sums, double* scratch, int planes, long long plane, long long plane_stride,
                   cudaStream_t stream) {
  UB_CHECK_ARG(src && sums && scratch && planes > 0 && plane > 0 && plane_stride >= plane, "zscore_sums: bad args");
  const int nblk = grid_for(plane, TPB * 8, UB_ZSCORE_BLOCKS);
  const dim3 grid(nblk, planes);
  if (src_dtype == 0) zscore_part_stats_kernel<uint8_t><<<grid, TPB, 0, stream>>>((const uint8_t*)src, scratch, plane, plane_stride);
  else if (src_dtype == 1) zscore_part_stats_kernel<uint16_t><<<grid, TPB, 0, stream>>>((const uint16_t*)src, scratch, plane, plane_stride);
  else if (src_dtype == 2) zscore_part_stats_kernel<float><<<grid, TPB, 0, stream>>>((const float*)src, scratch, plane, plane_stride);
  else UB_CHECK_ARG(false, "zscore_sums: bad src dtype %d", src_dtype);
  UB_LAUNCH_CHECK();
  zscore_fold_partials_kernel<<<planes, TPB, 0, stream>>>(scratch, sums, nblk);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_zscore_apply_sums(const void* src, int src_dtype, float* dst, const double* sums, double count, int planes, long long plane,
                         cudaStream_t stream) {
  UB_CHECK_ARG(src && dst && sums && count > 0 && planes > 0 && plane > 0, "zscore_apply_sums: bad args");
  const dim3 grid(grid_for(plane, TPB * 4, 4096), planes);
  if (src_dtype == 0) zscore_apply_sums_kernel<uint8_t><<<grid, TPB, 0, stream>>>((const uint8_t*)src, dst, sums, count, plane);
  else if (src_dtype == 1) zscore_apply_sums_kernel<uint16_t><<<grid, TPB, 0, stream>>>((const uint16_t*)src, dst, sums, count, plane);
  else if (src_dtype == 2) zscore_apply_sums_kernel<float><<<grid, TPB, 0, stream>>>((const float*)src, dst, sums, count, plane);
  else UB_CHECK_ARG(false, "zscore_apply_sums: bad src dtype %d", src_dtype);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

}  // extern "C"
