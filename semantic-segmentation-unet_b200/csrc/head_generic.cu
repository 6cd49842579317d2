// Class head for 8 < K <= UB_MAX_CLASSES_ANY classes (UNet/model.py:136-142, :211-215 with an arbitrary number_classes): the
// K-templated kernels of head.cu keep every class weight in registers, which stops at K = 8.  Here classes are spread over the
// LANES of a warp (class k = lane + 32 j), weights live in shared memory, and a warp reduces over classes with shuffles.  Same
// entry points, same buffers and partial-row layouts as head.cu (which dispatches here); HBM-bound only for small K -- at K = 32
// the 1x1 contraction is 2 x 64 x 32 flop per pixel on CUDA cores, ~0.5 ms per pass at 16 x 512 x 512.
#include "common.cuh"

namespace {

constexpr int TPB = 256;
constexpr int WARPS = TPB / 32;
constexpr int JMAX = (UB_MAX_CLASSES_ANY + 31) / 32;      // class chunks of 32 per lane

template <typename T>
__device__ __forceinline__ void ld8g(const T* p, float (&f)[8]);
template <>
__device__ __forceinline__ void ld8g<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
template <>
__device__ __forceinline__ void ld8g<float>(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// x tile [32 pixels][64] fp32 in shared memory (one 16-byte load per thread); pixels beyond P are zero
template <typename T>
__device__ __forceinline__ void load_x_tile(const T* __restrict__ x, long long px0, long long P, float (*xs)[64]) {
  const int p = threadIdx.x >> 3, sub = threadIdx.x & 7;
  float f[8] = {};
  if (px0 + p < P) ld8g<T>(x + (px0 + p) * 64 + sub * 8, f);
#pragma unroll
  for (int i = 0; i < 8; ++i) xs[p][sub * 8 + i] = f[i];
}

// ------------------------------------------------------------------ forward: a = relu(x . w^T + b) (+ statistics), or argmax / softmax
// MODE 0: training / plain forward (a_out, partial).  MODE 1: inference (y = a * scale + shift; argmax -> mask zone, optional softmax).
struct TileGeoG {
  int cy0, cy1, cx0, cx1, dst_y, dst_x;
};
template <typename T, int MODE>
__global__ void __launch_bounds__(TPB) head_fwd_g_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, int K,
                                                         int KP, long long P, float* __restrict__ a_out, float* __restrict__ partial,
                                                         const float* __restrict__ scale, const float* __restrict__ shift, int h, int wd,
                                                         const TileGeoG* __restrict__ geo, uint8_t* __restrict__ mask, long long mask_ld,
                                                         float* __restrict__ softmax_out) {
  extern __shared__ float smem[];
  float* wT = smem;                                         // [64][KP]
  float(*xs)[64] = reinterpret_cast<float(*)[64]>(smem + 64 * KP);      // [32][64]
  float* red = smem + 64 * KP + 32 * 64;                    // [WARPS][2][KP] (MODE 0)
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * KP; i += TPB) {
    const int c = i / KP, k = i % KP;
    wT[i] = k < K ? w[k * 64 + c] : 0.f;
  }
  float bk[JMAX], sc[JMAX], sf[JMAX], s[JMAX], q[JMAX];
  const int J = KP / 32;
#pragma unroll
  for (int j = 0; j < JMAX; ++j) {
    const int k = lane + 32 * j;
    bk[j] = (j < J && k < K) ? b[k] : 0.f;
    sc[j] = (MODE == 1 && j < J && k < K) ? scale[k] : 1.f;
    sf[j] = (MODE == 1 && j < J && k < K) ? shift[k] : 0.f;
    s[j] = q[j] = 0.f;
  }
  const int tile = MODE == 1 ? blockIdx.y : 0;
  const TileGeoG g = (MODE == 1 && geo) ? geo[tile] : TileGeoG{0, h, 0, wd, 0, 0};
  const T* xt = x + (long long)tile * P * 64;
  const long long iters = (P + 31) / 32;
  for (long long it = blockIdx.x; it < iters; it += gridDim.x) {
    __syncthreads();
    load_x_tile<T>(xt, it * 32, P, xs);
    __syncthreads();
    float y[4][JMAX];
#pragma unroll
    for (int j = 0; j < JMAX; ++j) {
      if (j < J) {
        float acc[4] = {bk[j], bk[j], bk[j], bk[j]};
        const float* wc = wT + lane + 32 * j;
#pragma unroll 8
        for (int c = 0; c < 64; ++c) {
          const float wv = wc[c * KP];
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u] = fmaf(xs[wrp * 4 + u][c], wv, acc[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) y[u][j] = fmaxf(acc[u], 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long px = it * 32 + wrp * 4 + u;
      if (px >= P) continue;                                // warp-uniform
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < JMAX; ++j) {
          const int k = lane + 32 * j;
          if (j < J && k < K) {
            a_out[px * K + k] = y[u][j];
            s[j] += y[u][j];
            q[j] = fmaf(y[u][j], y[u][j], q[j]);
          }
        }
      } else {
        float mx = -INFINITY;
        int am = 0;
#pragma unroll
        for (int j = 0; j < JMAX; ++j) {
          const int k = lane + 32 * j;
          if (j < J && k < K) {
            y[u][j] = fmaf(y[u][j], sc[j], sf[j]);
            if (y[u][j] > mx) {                              // ascending k within the lane: first maximum wins
              mx = y[u][j];
              am = k;
            }
          }
        }
        // argmax over lanes, lowest class index on ties (np.argmax)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float om = __shfl_xor_sync(0xffffffffu, mx, o);
          const int oa = __shfl_xor_sync(0xffffffffu, am, o);
          if (om > mx || (om == mx && oa < am)) {
            mx = om;
            am = oa;
          }
        }
        const int ty = (int)(px / wd), tx = (int)(px % wd);
        if (lane == 0 && mask && ty >= g.cy0 && ty < g.cy1 && tx >= g.cx0 && tx < g.cx1)
          mask[(long long)(g.dst_y + ty - g.cy0) * mask_ld + g.dst_x + tx - g.cx0] = (uint8_t)am;
        if (softmax_out) {
          float se = 0.f;
#pragma unroll
          for (int j = 0; j < JMAX; ++j) {
            const int k = lane + 32 * j;
            if (j < J && k < K) {
              y[u][j] = __expf(y[u][j] - mx);
              se += y[u][j];
            }
          }
          se = warp_sum(se);
          const float inv = 1.f / se;
#pragma unroll
          for (int j = 0; j < JMAX; ++j) {
            const int k = lane + 32 * j;
            if (j < J && k < K) softmax_out[((long long)tile * P + px) * K + k] = y[u][j] * inv;
          }
        }
      }
    }
  }
  if (MODE == 0 && partial) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < JMAX; ++j)
      if (j < J) {
        red[(wrp * 2 + 0) * KP + lane + 32 * j] = s[j];
        red[(wrp * 2 + 1) * KP + lane + 32 * j] = q[j];
      }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * K; i += TPB) {
      const int which = i / K, k = i % K;
      float t = 0.f;
#pragma unroll
      for (int v = 0; v < WARPS; ++v) t += red[(v * 2 + which) * KP + k];
      partial[((size_t)blockIdx.x * 2 + which) * K + k] = t;
    }
  }
}

// ------------------------------------------------------------------ loss: one warp per pixel, classes over lanes
__global__ void __launch_bounds__(TPB) head_loss_g_kernel(const float* __restrict__ a, const float* __restrict__ mean, const float* __restrict__ rstd,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          const uint8_t* __restrict__ labels, const float* __restrict__ class_w, float inv_denom,
                                                          float acc_scale, float smooth, float* __restrict__ softmax_out, float* __restrict__ dlogits,
                                                          float* __restrict__ partial, long long P, int K) {
  __shared__ float shl[WARPS], shc[WARPS];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int J = (K + 31) / 32;
  float sc[JMAX], sf[JMAX];
#pragma unroll
  for (int j = 0; j < JMAX; ++j) {
    const int k = lane + 32 * j;
    const bool ok = j < J && k < K;
    sc[j] = ok ? gamma[k] * rstd[k] : 0.f;
    sf[j] = ok ? beta[k] - mean[k] * sc[j] : 0.f;
  }
  float loss = 0.f, correct = 0.f;
  for (long long px = (long long)blockIdx.x * WARPS + wrp; px < P; px += (long long)gridDim.x * WARPS) {
    float y[JMAX];
    float mx = -INFINITY;
    int am = 0;
#pragma unroll
    for (int j = 0; j < JMAX; ++j) {
      const int k = lane + 32 * j;
      if (j < J && k < K) {
        y[j] = fmaf(__ldg(a + px * K + k), sc[j], sf[j]);
        if (y[j] > mx) {
          mx = y[j];
          am = k;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oa = __shfl_xor_sync(0xffffffffu, am, o);
      if (om > mx || (om == mx && oa < am)) {
        mx = om;
        am = oa;
      }
    }
    float se = 0.f, sy = 0.f;
#pragma unroll
    for (int j = 0; j < JMAX; ++j) {
      const int k = lane + 32 * j;
      if (j < J && k < K) {
        sy += y[j] - mx;
        y[j] = __expf(y[j] - mx);
        se += y[j];
      }
    }
    se = warp_sum(se);
    const float inv = 1.f / se;
    const int lab = labels ? (int)labels[px] : 0;
    const float cw = class_w ? class_w[lab] : 1.f;
    const float t_on = 1.f - smooth + smooth / (float)K, t_off = smooth / (float)K;      // label smoothing, see head.cu
    float pl = 0.f;
#pragma unroll
    for (int j = 0; j < JMAX; ++j) {
      const int k = lane + 32 * j;
      if (j < J && k < K) {
        const float p = y[j] * inv;
        if (k == lab) pl = p;
        if (softmax_out) softmax_out[px * K + k] = p;
        if (dlogits) dlogits[px * K + k] = (p - (k == lab ? t_on : t_off)) * cw * inv_denom;
      }
    }
    if (labels) {
      pl = warp_sum(pl);                                     // exactly one lane holds the label's probability
      if (smooth != 0.f) sy = warp_sum(sy);
      if (lane == 0) {
        float l = -__logf(fmaxf(pl, 1e-37f)) * (1.f - smooth);
        if (smooth != 0.f) l -= t_off * (sy - (float)K * __logf(se));
        loss += l * cw;
        correct += (am == lab) ? 1.f : 0.f;
      }
    }
  }
  if (partial) {
    if (lane == 0) {
      shl[wrp] = loss;
      shc[wrp] = correct;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float tl = 0.f, tc = 0.f;
#pragma unroll
      for (int v = 0; v < WARPS; ++v) {
        tl += shl[v];
        tc += shc[v];
      }
      partial[(size_t)blockIdx.x * 2 + 0] = tl * inv_denom;
      partial[(size_t)blockIdx.x * 2 + 1] = tc * acc_scale;
    }
  }
}

// ------------------------------------------------------------------ backward sums: partial[row][0][k] = sum dy_k, [row][1][k] = sum dy_k xhat_k
__global__ void __launch_bounds__(TPB) head_bwd_reduce_g_kernel(const float* __restrict__ dy, const float* __restrict__ a, const float* __restrict__ mean,
                                                                const float* __restrict__ rstd, float* __restrict__ partial, long long P, int K) {
  extern __shared__ float red[];                             // [WARPS][2][KP]
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int J = (K + 31) / 32, KP = J * 32;
  float mu[JMAX], rs[JMAX], s[JMAX], q[JMAX];
#pragma unroll
  for (int j = 0; j < JMAX; ++j) {
    const int k = lane + 32 * j;
    const bool ok = j < J && k < K;
    mu[j] = ok ? mean[k] : 0.f;
    rs[j] = ok ? rstd[k] : 0.f;
    s[j] = q[j] = 0.f;
  }
  for (long long px = (long long)blockIdx.x * WARPS + wrp; px < P; px += (long long)gridDim.x * WARPS) {
#pragma unroll
    for (int j = 0; j < JMAX; ++j) {
      const int k = lane + 32 * j;
      if (j < J && k < K) {
        const float d = __ldg(dy + px * K + k);
        s[j] += d;
        q[j] = fmaf(d, (__ldg(a + px * K + k) - mu[j]) * rs[j], q[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < JMAX; ++j)
    if (j < J) {
      red[(wrp * 2 + 0) * KP + lane + 32 * j] = s[j];
      red[(wrp * 2 + 1) * KP + lane + 32 * j] = q[j];
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * K; i += TPB) {
    const int which = i / K, k = i % K;
    float t = 0.f;
#pragma unroll
    for (int v = 0; v < WARPS; ++v) t += red[(v * 2 + which) * KP + k];
    partial[((size_t)blockIdx.x * 2 + which) * K + k] = t;
  }
}

// dz_k = gamma_k rstd_k (dy_k - dbeta_k / P - xhat_k dgamma_k / P) [a_k > 0] for a 32-pixel tile -> dzs[32][KP] (class lanes, 4 pixels per warp)
__device__ __forceinline__ void dz_tile(const float* __restrict__ dy, const float* __restrict__ a, long long px0, long long P, int K, int KP,
                                        const float* __restrict__ cst /*[5][KP]: mean, rstd, gamma*rstd, dbeta/P, dgamma/P*/, float* dzs) {
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  for (int u = 0; u < 4; ++u) {
    const int p = wrp * 4 + u;
    const long long px = px0 + p;
    for (int k = lane; k < KP; k += 32) {
      float dz = 0.f;
      if (px < P && k < K) {
        const float av = __ldg(a + px * K + k);
        const float xh = (av - cst[k]) * cst[KP + k];
        dz = cst[2 * KP + k] * (__ldg(dy + px * K + k) - cst[3 * KP + k] - xh * cst[4 * KP + k]);
        if (!(av > 0.f)) dz = 0.f;
      }
      dzs[p * KP + k] = dz;
    }
  }
}

__device__ __forceinline__ void load_consts(float* cst, int K, int KP, const float* mean, const float* rstd, const float* gamma, const float* dbeta,
                                            const float* dgamma, float invP) {
  for (int k = threadIdx.x; k < KP; k += TPB) {
    const bool ok = k < K;
    cst[k] = ok ? mean[k] : 0.f;
    cst[KP + k] = ok ? rstd[k] : 0.f;
    cst[2 * KP + k] = ok ? gamma[k] * rstd[k] : 0.f;
    cst[3 * KP + k] = ok ? dbeta[k] * invP : 0.f;
    cst[4 * KP + k] = ok ? dgamma[k] * invP : 0.f;
  }
}

// ------------------------------------------------------------------ backward apply, part 1: dx[p][c] = sum_k w[k][c] dz_k, db partials
template <typename T>
__global__ void __launch_bounds__(TPB) head_bwd_dx_g_kernel(const float* __restrict__ dy, const float* __restrict__ a, const float* __restrict__ w,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ dbeta,
                                                            const float* __restrict__ dgamma, T* __restrict__ dx, float* __restrict__ partial,
                                                            long long P, int K, int KP) {
  extern __shared__ float smem[];
  float* ws = smem;                      // [KP][64]
  float* dzs = ws + KP * 64;             // [32][KP]
  float* cst = dzs + 32 * KP;            // [5][KP]
  float* dbs = cst + 5 * KP;             // [KP] block total of dz per class
  for (int i = threadIdx.x; i < KP * 64; i += TPB) ws[i] = (i / 64) < K ? w[i] : 0.f;
  load_consts(cst, K, KP, mean, rstd, gamma, dbeta, dgamma, 1.f / (float)P);
  for (int k = threadIdx.x; k < KP; k += TPB) dbs[k] = 0.f;
  const int c = threadIdx.x & 63, pg = threadIdx.x >> 6;     // dx mapping: channel c, pixels pg * 8 .. + 7 of the tile
  const long long iters = (P + 31) / 32;
  for (long long it = blockIdx.x; it < iters; it += gridDim.x) {
    __syncthreads();
    dz_tile(dy, a, it * 32, P, K, KP, cst, dzs);
    __syncthreads();
    float o[8] = {};
    for (int k = 0; k < K; ++k) {
      const float wv = ws[k * 64 + c];
#pragma unroll
      for (int u = 0; u < 8; ++u) o[u] = fmaf(dzs[(pg * 8 + u) * KP + k], wv, o[u]);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const long long px = it * 32 + pg * 8 + u;
      if (px < P && dx) dx[px * 64 + c] = (T)o[u];
    }
    // per-class sum of dz over the tile's pixels (deterministic: one thread per class, fixed pixel order)
    for (int k = threadIdx.x; k < K; k += TPB) {
      float t = 0.f;
#pragma unroll 8
      for (int p = 0; p < 32; ++p) t += dzs[p * KP + k];
      dbs[k] += t;
    }
  }
  __syncthreads();
  const int ncomp = K * 64 + K;
  for (int k = threadIdx.x; k < K; k += TPB) partial[(size_t)blockIdx.x * ncomp + K * 64 + k] = dbs[k];
}

// ------------------------------------------------------------------ backward apply, part 2: dW[k][c] = sum_p dz[p][k] x[p][c]
// thread = (channel c, class group kq of 4): classes kq + 4 j, j < KJ accumulators in registers; dz is recomputed from dy / a per tile
template <typename T, int KJ>
__global__ void __launch_bounds__(TPB) head_bwd_dw_g_kernel(const float* __restrict__ dy, const float* __restrict__ a, const T* __restrict__ x,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ dbeta,
                                                            const float* __restrict__ dgamma, float* __restrict__ partial, long long P, int K,
                                                            int KP) {
  extern __shared__ float smem[];
  float(*xs)[64] = reinterpret_cast<float(*)[64]>(smem);     // [32][64]
  float* dzs = smem + 32 * 64;                               // [32][KP]
  float* cst = dzs + 32 * KP;                                // [5][KP]
  load_consts(cst, K, KP, mean, rstd, gamma, dbeta, dgamma, 1.f / (float)P);
  const int c = threadIdx.x & 63, kq = threadIdx.x >> 6;
  float acc[KJ];
#pragma unroll
  for (int j = 0; j < KJ; ++j) acc[j] = 0.f;
  const long long iters = (P + 31) / 32;
  for (long long it = blockIdx.x; it < iters; it += gridDim.x) {
    __syncthreads();
    load_x_tile<T>(x, it * 32, P, xs);
    dz_tile(dy, a, it * 32, P, K, KP, cst, dzs);
    __syncthreads();
    for (int p = 0; p < 32; ++p) {
      const float xv = xs[p][c];
      const float* dp = dzs + p * KP + kq;
#pragma unroll
      for (int j = 0; j < KJ; ++j) acc[j] = fmaf(dp[4 * j], xv, acc[j]);      // 4 * j < KP: KJ = KP / 4
    }
  }
  const int ncomp = K * 64 + K;
#pragma unroll
  for (int j = 0; j < KJ; ++j) {
    const int k = kq + 4 * j;
    if (k < K) partial[(size_t)blockIdx.x * ncomp + k * 64 + c] = acc[j];
  }
}

inline int grid_rows(long long work_items, int per_block) {
  long long g = (work_items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > UB_STATS_ROWS) g = UB_STATS_ROWS;
  return (int)g;
}

template <typename K>
int set_smem(K kern, size_t bytes) {
  if (bytes > 48 * 1024) UB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return UB_OK;
}

}  // namespace

// ---- called by the extern "C" entry points of head.cu for K > UB_MAX_CLASSES ---------------------------------------------------------
int ubg_head_fwd(const void* x, const float* w, const float* b, float* a_out, float* partial, long long P, int K, int dtype, cudaStream_t stream) {
  const int KP = (K + 31) / 32 * 32;
  const size_t smem = sizeof(float) * (64 * KP + 32 * 64 + WARPS * 2 * KP);
  const int grid = grid_rows(P, 32 * 8);
  int rc;
  if (partial) UB_CUDA(cudaMemsetAsync(partial, 0, sizeof(float) * UB_STATS_ROWS * 2 * K, stream));
  if (dtype == UB_BF16) {
    if ((rc = set_smem(head_fwd_g_kernel<__nv_bfloat16, 0>, smem))) return rc;
    head_fwd_g_kernel<__nv_bfloat16, 0><<<grid, TPB, smem, stream>>>((const __nv_bfloat16*)x, w, b, K, KP, P, a_out, partial, nullptr, nullptr, 0, 0,
                                                                   nullptr, nullptr, 0, nullptr);
  } else {
    if ((rc = set_smem(head_fwd_g_kernel<float, 0>, smem))) return rc;
    head_fwd_g_kernel<float, 0><<<grid, TPB, smem, stream>>>((const float*)x, w, b, K, KP, P, a_out, partial, nullptr, nullptr, 0, 0, nullptr, nullptr,
                                                           0, nullptr);
  }
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ubg_head_argmax(const void* x, const float* w, const float* b, const float* scale, const float* shift, int K, int ntiles, int h, int wd,
                    const int* geo, unsigned char* mask, long long mask_ld, float* softmax_out, int dtype, cudaStream_t stream) {
  const int KP = (K + 31) / 32 * 32;
  const size_t smem = sizeof(float) * (64 * KP + 32 * 64 + WARPS * 2 * KP);
  const long long P = (long long)h * wd;
  int gx = grid_rows(P, 32 * 8);
  const int cap = ub_num_sms() * 8 / (ntiles < 8 ? ntiles : 8);
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  dim3 grid(gx, ntiles);
  int rc;
  if (dtype == UB_BF16) {
    if ((rc = set_smem(head_fwd_g_kernel<__nv_bfloat16, 1>, smem))) return rc;
    head_fwd_g_kernel<__nv_bfloat16, 1><<<grid, TPB, smem, stream>>>((const __nv_bfloat16*)x, w, b, K, KP, P, nullptr, nullptr, scale, shift, h, wd,
                                                                   reinterpret_cast<const TileGeoG*>(geo), mask, mask_ld, softmax_out);
  } else {
    if ((rc = set_smem(head_fwd_g_kernel<float, 1>, smem))) return rc;
    head_fwd_g_kernel<float, 1><<<grid, TPB, smem, stream>>>((const float*)x, w, b, K, KP, P, nullptr, nullptr, scale, shift, h, wd,
                                                           reinterpret_cast<const TileGeoG*>(geo), mask, mask_ld, softmax_out);
  }
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ubg_head_loss(const float* a, const float* mean, const float* rstd, const float* gamma, const float* beta, const unsigned char* labels,
                  const float* class_w, float inv_denom, float acc_scale, float smooth, float* softmax_out, float* dlogits, float* partial, long long P,
                  int K, cudaStream_t stream) {
  if (partial) UB_CUDA(cudaMemsetAsync(partial, 0, sizeof(float) * UB_STATS_ROWS * 2, stream));
  head_loss_g_kernel<<<grid_rows(P, WARPS * 16), TPB, 0, stream>>>(a, mean, rstd, gamma, beta, labels, class_w, inv_denom, acc_scale, smooth, softmax_out,
                                                                  dlogits, partial, P, K);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ubg_head_bwd_reduce(const float* dy, const float* a, const float* mean, const float* rstd, float* partial, long long P, int K,
                        cudaStream_t stream) {
  const int KP = (K + 31) / 32 * 32;
  UB_CUDA(cudaMemsetAsync(partial, 0, sizeof(float) * UB_STATS_ROWS * 2 * K, stream));
  head_bwd_reduce_g_kernel<<<grid_rows(P, WARPS * 16), TPB, sizeof(float) * WARPS * 2 * KP, stream>>>(dy, a, mean, rstd, partial, P, K);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ubg_head_bwd_apply(const float* dy, const float* a, const void* x, const float* w, const float* mean, const float* rstd, const float* gamma,
                       const float* dbeta, const float* dgamma, void* dx, float* partial, long long P, int K, int dtype, cudaStream_t stream) {
  const int KP = (K + 31) / 32 * 32;
  const int grid = grid_rows(P, 32 * 8);
  const size_t ncomp = (size_t)K * 64 + K;
  UB_CUDA(cudaMemsetAsync(partial, 0, sizeof(float) * UB_STATS_ROWS * ncomp, stream));
  const size_t smem_dx = sizeof(float) * ((size_t)KP * 64 + 32 * KP + 5 * KP + KP);
  const size_t smem_dw = sizeof(float) * ((size_t)32 * 64 + 32 * KP + 5 * KP);
  int rc;
#define UBG_DW(TT, KJ)                                                                                                               \
  do {                                                                                                                               \
    if ((rc = set_smem(head_bwd_dw_g_kernel<TT, KJ>, smem_dw))) return rc;                                                           \
    head_bwd_dw_g_kernel<TT, KJ><<<grid, TPB, smem_dw, stream>>>(dy, a, (const TT*)x, mean, rstd, gamma, dbeta, dgamma, partial, P, K, KP); \
  } while (0)
#define UBG_DW_ALL(TT)                    \
  do {                                    \
    if (KP <= 32) UBG_DW(TT, 8);          \
    else if (KP <= 64) UBG_DW(TT, 16);    \
    else if (KP <= 128) UBG_DW(TT, 32);   \
    else UBG_DW(TT, 64);                  \
  } while (0)
  if (dtype == UB_BF16) {
    if ((rc = set_smem(head_bwd_dx_g_kernel<__nv_bfloat16>, smem_dx))) return rc;
    head_bwd_dx_g_kernel<__nv_bfloat16><<<grid, TPB, smem_dx, stream>>>(dy, a, w, mean, rstd, gamma, dbeta, dgamma, (__nv_bfloat16*)dx, partial, P, K, KP);
    UB_LAUNCH_CHECK();
    UBG_DW_ALL(__nv_bfloat16);
  } else {
    if ((rc = set_smem(head_bwd_dx_g_kernel<float>, smem_dx))) return rc;
    head_bwd_dx_g_kernel<float><<<grid, TPB, smem_dx, stream>>>(dy, a, w, mean, rstd, gamma, dbeta, dgamma, (float*)dx, partial, P, K, KP);
    UB_LAUNCH_CHECK();
    UBG_DW_ALL(float);
  }
#undef UBG_DW_ALL
#undef UBG_DW
  UB_LAUNCH_CHECK();
  return UB_OK;
}
