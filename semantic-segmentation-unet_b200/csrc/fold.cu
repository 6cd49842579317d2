// BatchNorm folded into the CONSUMER convolution (training forward), DESIGN.md "the plan for it":
//   y = s * a + t per input channel (s = gamma * rstd, t = beta - mean * s of the PRODUCER's BatchNorm, UNet/model.py:36),
//   conv(y) = conv_{W * s}(a) + sum over the taps that fall inside the image of (W . t)      ('same' padding pads y, not a)
// so the consumer reads the pre-BatchNorm activation `a` and the y tensor is never written.  Small helper kernels:
//   ub_fold_conv3_weights  W' = bf16(W * s[ci]), the 9-case border bias table, and s / t for the backward fix-up
//   ub_border_sums         Sdz[tap][co] = sum of dz over the pixels whose tap neighbour is inside the image
//   ub_wgrad_fold_fix      dW = s[ci] * dW_a + t[ci] * Sdz[tap][co]      (dW_a = weight gradient computed on `a`)
// The algebra is proven in fp64 in oracle/unet_numpy.py (fold_weights, border_case_bias, border_sums, conv_wgrad_folded).
//   ub_fold_head_weights / ub_head_wgrad_fold_fix   the same for the 1x1 head (no padding: a single bias vector)
// Parity: tests/kernel_cases.py (fold_weights_*, conv_fwd_folded_*, wgrad_folded_*, head_fold_*) and the whole-graph cases under
// UB_FOLD_BN=1 (tests/graph_cases.py fold_*), green on B200.
#include "common.cuh"

namespace {

constexpr int TPB = 256;

__device__ __forceinline__ void channel_affine(int ci, int C0, const float* mean0, const float* rstd0, const float* gamma0, const float* beta0,
                                               const float* mean1, const float* rstd1, const float* gamma1, const float* beta1, float& s,
                                               float& t) {
  const bool first = ci < C0;
  const float* mean = first ? mean0 : mean1;
  if (!mean) {          // this source is a real (already normalised) tensor: identity
    s = 1.f;
    t = 0.f;
    return;
  }
  const int c = first ? ci : ci - C0;
  s = (first ? gamma0 : gamma1)[c] * (first ? rstd0 : rstd1)[c];
  t = (first ? beta0 : beta1)[c] - mean[c] * s;
}

// one block per output channel: w [Cout][9][Cin] fp32 -> w_out bf16, Tt[tap] = sum_ci w * t (fp64), bias9[case][co]
__global__ void __launch_bounds__(TPB) fold_conv3_kernel(const float* __restrict__ w, int Cout, int C0, const float* mean0, const float* rstd0,
                                                         const float* gamma0, const float* beta0, int C1, const float* mean1,
                                                         const float* rstd1, const float* gamma1, const float* beta1,
                                                         const float* __restrict__ bias, __nv_bfloat16* __restrict__ w_out,
                                                         float* __restrict__ bias9, float* __restrict__ scale_out,
                                                         float* __restrict__ shift_out) {
  const int co = blockIdx.x;
  const int Cin = C0 + C1;
  double tt[9] = {};
  for (int ci = threadIdx.x; ci < Cin; ci += TPB) {
    float s, t;
    channel_affine(ci, C0, mean0, rstd0, gamma0, beta0, mean1, rstd1, gamma1, beta1, s, t);
    if (co == 0) {
      scale_out[ci] = s;
      shift_out[ci] = t;
    }
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const size_t i = ((size_t)co * 9 + tap) * Cin + ci;
      const float wv = w[i];
      w_out[i] = __float2bfloat16(wv * s);
      tt[tap] += (double)wv * (double)t;
    }
  }
  __shared__ double sh[9][TPB / 32];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    double v = tt[tap];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[tap][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    // border case k = row case * 3 + column case: 0 = first row / column (tap offset -1 is outside), 1 = interior, 2 = last
    const int rc = threadIdx.x / 3, cc = threadIdx.x % 3;
    double acc = bias ? (double)bias[co] : 0.0;
    for (int dy = 0; dy < 3; ++dy) {
      if ((rc == 0 && dy == 0) || (rc == 2 && dy == 2)) continue;
      for (int dx = 0; dx < 3; ++dx) {
        if ((cc == 0 && dx == 0) || (cc == 2 && dx == 2)) continue;
        double tsum = 0.0;
        for (int wv = 0; wv < TPB / 32; ++wv) tsum += sh[dy * 3 + dx][wv];
        acc += tsum;
      }
    }
    bias9[(size_t)threadIdx.x * Cout + co] = (float)acc;
  }
}

// raw[chunk][s][c]: s = 0 top row, 1 bottom row, 2 left column, 3 right column, 4..7 corners (tl, tr, bl, br); NHWC dz.
// The strips hold N * (2 W + 2 H) pixels of C channels: a few MB, but spread over the whole tensor -- the work is split over
// UB_BORDER_CHUNKS blocks per strip so that enough independent loads are in flight (the first version walked a strip with 4 pixel
// lanes per 64 channels: 2048 dependent iterations, 0.5 ms per launch at 16 x 512 x 512).
template <typename T>
__global__ void __launch_bounds__(TPB) border_strips_kernel(const T* __restrict__ dz, float* __restrict__ raw, int N, int H, int W, int C) {
  const int s = blockIdx.y;
  const int c = blockIdx.x * 64 + (threadIdx.x & 63);
  const int g = threadIdx.x >> 6;          // 4 pixel lanes
  const long long count = s < 2 ? (long long)N * W : (s < 4 ? (long long)N * H : N);
  const long long lanes = 4ll * gridDim.z;
  float acc = 0.f;
  if (c < C) {
#pragma unroll 4
    for (long long p = (long long)blockIdx.z * 4 + g; p < count; p += lanes) {
      int n, h, w;
      if (s < 2) {
        n = (int)(p / W);
        w = (int)(p % W);
        h = s == 0 ? 0 : H - 1;
      } else if (s < 4) {
        n = (int)(p / H);
        h = (int)(p % H);
        w = s == 2 ? 0 : W - 1;
      } else {
        n = (int)p;
        h = (s == 4 || s == 5) ? 0 : H - 1;
        w = (s == 4 || s == 6) ? 0 : W - 1;
      }
      acc += (float)dz[(((long long)n * H + h) * W + w) * C + c];
    }
  }
  __shared__ float sh[4][64];
  sh[g][threadIdx.x & 63] = acc;
  __syncthreads();
  if (g == 0 && c < C)
    raw[((size_t)blockIdx.z * 8 + s) * C + c] = (sh[0][threadIdx.x] + sh[1][threadIdx.x]) + (sh[2][threadIdx.x] + sh[3][threadIdx.x]);
}

// sdz[tap][c] = total - (row strip excluded by the tap) - (column strip excluded by the tap) + (their common corner).
// One block = 32 channels x 8 strips; the chunk partials of a (strip, channel) are summed by one thread with all loads of an
// 8-chunk group issued before the first add (64 dependent L2 round trips per thread made the first version 88 us per launch).
__global__ void __launch_bounds__(256) border_finish_kernel(const float* __restrict__ raw, const float* __restrict__ total, float* __restrict__ sdz,
                                                            int C, int chunks) {
  __shared__ double r[8][32];
  const int lane = threadIdx.x & 31, s = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double acc = 0.0;
  if (c < C) {
    for (int k0 = 0; k0 < chunks; k0 += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (k0 + u < chunks) ? __ldg(raw + ((size_t)(k0 + u) * 8 + s) * C + c) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += (double)v[u];
    }
  }
  r[s][lane] = acc;
  __syncthreads();
  if (c >= C) return;
  const double tot = (double)total[c];
  for (int tap = s; tap < 9; tap += 8) {
    const int dy = tap / 3, dx = tap % 3;
    double v = tot;
    if (dy == 0) v -= r[0][lane];
    if (dy == 2) v -= r[1][lane];
    if (dx == 0) v -= r[2][lane];
    if (dx == 2) v -= r[3][lane];
    if (dy != 1 && dx != 1) v += r[4 + (dy == 2 ? 2 : 0) + (dx == 2 ? 1 : 0)][lane];
    sdz[(size_t)tap * C + c] = (float)v;
  }
}

__global__ void __launch_bounds__(TPB) wgrad_fold_fix_kernel(float* __restrict__ dw, const float* __restrict__ scale,
                                                             const float* __restrict__ shift, const float* __restrict__ sdz, int Cout, int Cin) {
  const long long n = (long long)Cout * 9 * Cin;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < n; i += (long long)gridDim.x * TPB) {
    const int ci = (int)(i % Cin);
    const long long r = i / Cin;
    const int tap = (int)(r % 9), co = (int)(r / 9);
    dw[i] = fmaf(scale[ci], dw[i], shift[ci] * sdz[(size_t)tap * Cout + co]);
  }
}

// Backward sums of the BatchNorm that FEEDS a convolution, from that convolution's weight gradient (oracle/unet_numpy.py
// bn_bwd_sums_from_wgrad): with dy = dgrad(dz, W) the gradient w.r.t. the BatchNorm output,
//   dbeta[c]  = sum_p dy[p][c]            = sum_{co, tap} W[co][tap][c] * Sdz[tap][co]
//   dgamma[c] = sum_p dy[p][c] xhat[p][c] = rstd[c] * sum_{co, tap} W[co][tap][c] * (dW_a[co][tap][c] - mean[c] * Sdz[tap][co])
// dW_a = the consumer's weight gradient computed on the pre-BatchNorm activation `a` (before ub_wgrad_fold_fix), Sdz = its border
// sums (taps == 1: Sdz = the bias gradient).  Replaces a read of dy and `a` (two full tensors) by a read of two weight-sized ones.
// Block = 32 channels x 8 row groups; rows (co, tap) are strided over the groups, fp64 accumulation, fixed combination order.
template <typename TW>
__global__ void __launch_bounds__(256) bn_sums_wgrad_kernel(const TW* __restrict__ w, const float* __restrict__ dw, const float* __restrict__ sdz,
                                                           int Cout, int taps, int Cin, int c_begin, int c_count,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           float* __restrict__ dbeta, float* __restrict__ dgamma) {
  __shared__ double sh[2][8][32];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double sb = 0.0, sg = 0.0;
  if (c < c_count) {
    const float mu = mean[c];
    const int rows = Cout * taps;
    for (int r0 = grp; r0 < rows; r0 += 8 * 4) {
      float wv[4], dv[4], sv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u * 8;
        if (r < rows) {
          const size_t i = (size_t)r * Cin + c_begin + c;
          wv[u] = (float)w[i];
          dv[u] = __ldg(dw + i);
          sv[u] = __ldg(sdz + (size_t)(r % taps) * Cout + r / taps);
        } else {
          wv[u] = dv[u] = sv[u] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        sb += (double)wv[u] * (double)sv[u];
        sg += (double)wv[u] * ((double)dv[u] - (double)mu * (double)sv[u]);
      }
    }
  }
  sh[0][grp][lane] = sb;
  sh[1][grp][lane] = sg;
  __syncthreads();
  if (grp == 0 && c < c_count) {
    double tb = 0.0, tg = 0.0;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      tb += sh[0][g][lane];
      tg += sh[1][g][lane];
    }
    dbeta[c] = (float)tb;
    dgamma[c] = (float)(tg * (double)rstd[c]);
  }
}

// 1x1 head (no padding: one bias): w' = w * s[c], b' = b + sum_c w * t[c]; one block, K <= UB_MAX_CLASSES_ANY
__global__ void fold_head_kernel(const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ mean,
                                 const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float* __restrict__ w_out, float* __restrict__ bias_out, float* __restrict__ scale_out,
                                 float* __restrict__ shift_out, int K) {
  __shared__ double sh[64];
  const int c = threadIdx.x;          // 64 threads = 64 input channels
  const float s = gamma[c] * rstd[c];
  const float t = beta[c] - mean[c] * s;
  scale_out[c] = s;
  shift_out[c] = t;
  for (int k = 0; k < K; ++k) {
    const float wv = w[k * 64 + c];
    w_out[k * 64 + c] = wv * s;
    __syncthreads();
    sh[c] = (double)wv * (double)t;
    __syncthreads();
    if (c == 0) {
      double acc = bias ? (double)bias[k] : 0.0;
      for (int i = 0; i < 64; ++i) acc += sh[i];
      bias_out[k] = (float)acc;
    }
  }
}

// dW[k][c] = s[c] * dW_a[k][c] + t[c] * db[k]
__global__ void head_wgrad_fold_fix_kernel(float* __restrict__ dw, const float* __restrict__ db, const float* __restrict__ scale,
                                           const float* __restrict__ shift, int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K * 64) dw[i] = fmaf(scale[i & 63], dw[i], shift[i & 63] * db[i >> 6]);
}

}  // namespace

extern "C" {

int ub_fold_head_weights(const float* w, const float* bias, const float* mean, const float* rstd, const float* gamma, const float* beta,
                         float* w_out, float* bias_out, float* scale_out, float* shift_out, int K, cudaStream_t stream) {
  UB_CHECK_ARG(w && mean && rstd && gamma && beta && w_out && bias_out && scale_out && shift_out && K >= 1 && K <= UB_MAX_CLASSES_ANY,
               "fold_head_weights: bad args");
  fold_head_kernel<<<1, 64, 0, stream>>>(w, bias, mean, rstd, gamma, beta, w_out, bias_out, scale_out, shift_out, K);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_head_wgrad_fold_fix(float* dw, const float* db, const float* scale, const float* shift, int K, cudaStream_t stream) {
  UB_CHECK_ARG(dw && db && scale && shift && K >= 1 && K <= UB_MAX_CLASSES_ANY, "head_wgrad_fold_fix: bad args");
  head_wgrad_fold_fix_kernel<<<(K * 64 + 127) / 128, 128, 0, stream>>>(dw, db, scale, shift, K);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_fold_conv3_weights(const float* w, int Cout, int C0, const float* mean0, const float* rstd0, const float* gamma0, const float* beta0,
                          int C1, const float* mean1, const float* rstd1, const float* gamma1, const float* beta1, const float* bias,
                          void* w_out, float* bias9, float* scale_out, float* shift_out, cudaStream_t stream) {
  UB_CHECK_ARG(w && w_out && bias9 && scale_out && shift_out && Cout > 0 && C0 > 0 && C1 >= 0, "fold_conv3_weights: bad args");
  UB_CHECK_ARG(!mean0 || (rstd0 && gamma0 && beta0), "fold_conv3_weights: source 0 needs mean, rstd, gamma and beta");
  UB_CHECK_ARG(!mean1 || (rstd1 && gamma1 && beta1 && C1 > 0), "fold_conv3_weights: source 1 needs mean, rstd, gamma and beta");
  fold_conv3_kernel<<<Cout, TPB, 0, stream>>>(w, Cout, C0, mean0, rstd0, gamma0, beta0, C1, mean1, rstd1, gamma1, beta1, bias,
                                             (__nv_bfloat16*)w_out, bias9, scale_out, shift_out);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_border_sums(const void* dz, const float* total, float* sdz, float* scratch, int N, int H, int W, int C, int dtype, cudaStream_t stream) {
  UB_CHECK_ARG(dz && total && sdz && scratch && N > 0 && H >= 2 && W >= 2 && C > 0,
               "border_sums: bad args (H, W >= 2; scratch = UB_BORDER_CHUNKS * 8 * C floats)");
  long long longest = (long long)N * (H > W ? H : W);
  int chunks = (int)((longest + 63) / 64);           // >= 16 pixels per lane where the strip is long enough
  if (chunks > UB_BORDER_CHUNKS) chunks = UB_BORDER_CHUNKS;
  if (chunks < 1) chunks = 1;
  const dim3 grid((C + 63) / 64, 8, chunks);
  if (dtype == UB_BF16) border_strips_kernel<__nv_bfloat16><<<grid, TPB, 0, stream>>>((const __nv_bfloat16*)dz, scratch, N, H, W, C);
  else if (dtype == UB_F32) border_strips_kernel<float><<<grid, TPB, 0, stream>>>((const float*)dz, scratch, N, H, W, C);
  else UB_CHECK_ARG(false, "border_sums: bad dtype %d", dtype);
  UB_LAUNCH_CHECK();
  border_finish_kernel<<<(C + 31) / 32, 256, 0, stream>>>(scratch, total, sdz, C, chunks);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_bn_bwd_sums_wgrad(const void* w, int w_dtype, const float* dw_a, const float* sdz, int Cout, int taps, int Cin, int c_begin, int c_count,
                         const float* mean, const float* rstd, float* dbeta, float* dgamma, cudaStream_t stream) {
  UB_CHECK_ARG(w && dw_a && sdz && mean && rstd && dbeta && dgamma, "bn_bwd_sums_wgrad: null pointer");
  UB_CHECK_ARG(Cout > 0 && taps > 0 && Cin > 0 && c_begin >= 0 && c_count > 0 && c_begin + c_count <= Cin, "bn_bwd_sums_wgrad: bad channel range");
  const int grid = (c_count + 31) / 32;
  if (w_dtype == UB_BF16)
    bn_sums_wgrad_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)w, dw_a, sdz, Cout, taps, Cin, c_begin, c_count, mean, rstd, dbeta, dgamma);
  else if (w_dtype == UB_F32)
    bn_sums_wgrad_kernel<float><<<grid, 256, 0, stream>>>((const float*)w, dw_a, sdz, Cout, taps, Cin, c_begin, c_count, mean, rstd, dbeta, dgamma);
  else
    UB_CHECK_ARG(false, "bn_bwd_sums_wgrad: bad weight dtype %d", w_dtype);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_wgrad_fold_fix(float* dw, const float* scale, const float* shift, const float* sdz, int Cout, int Cin, cudaStream_t stream) {
  UB_CHECK_ARG(dw && scale && shift && sdz && Cout > 0 && Cin > 0, "wgrad_fold_fix: bad args");
  const long long n = (long long)Cout * 9 * Cin;
  long long blocks = (n + TPB - 1) / TPB;
  if (blocks > (long long)ub_num_sms() * 16) blocks = (long long)ub_num_sms() * 16;
  wgrad_fold_fix_kernel<<<(int)blocks, TPB, 0, stream>>>(dw, scale, shift, sdz, Cout, Cin);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

}  // extern "C"
