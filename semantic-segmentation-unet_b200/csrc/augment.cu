// Training-time image augmentation on the device (UNet/augment.py:19-174, called per example from
// UNet/imagereader.py:283-294): affine warps with mirror boundary, additive Gaussian noise, Gaussian blur, additive
// intensity shift -- on whole raw-pixel batches that are already in HBM, between the H2D copy and the z-score.
// All of it is gather / elementwise work: HBM- and L2-bound, one thread per output element, fp64 coordinate math
// (skimage computes warps in double), planes laid out NCHW like the reader ships them.
#include "common.cuh"

namespace {

constexpr int TPB = 256;
constexpr int MM_BLOCKS = 64;          // min/max partial blocks per image
constexpr int MAX_RADIUS = 32;         // Gaussian kernel radius = int(4 sigma + 0.5); the reference's sigma <= 2 gives 8

// skimage _shared/interpolation.pxd coord_map, mode 'R' (numpy.pad 'reflect': mirror WITHOUT repeating the edge sample)
__device__ __forceinline__ long long reflect_nodup(long long dim, long long coord) {
  const long long cmax = dim - 1;
  if (dim == 1) return 0;
  if (coord < 0) {
    const long long n = -coord;
    return ((n / cmax) & 1) ? cmax - (n % cmax) : n % cmax;
  }
  if (coord > cmax) return ((coord / cmax) & 1) ? cmax - (coord % cmax) : coord % cmax;
  return coord;
}

// scipy.ndimage mode 'reflect' (d c b a | a b c d | d c b a: the edge sample IS repeated)
__device__ __forceinline__ int reflect_dup(int dim, int i) {
  while (i < 0 || i >= dim) i = i < 0 ? -i - 1 : 2 * dim - 1 - i;
  return i;
}

template <typename T>
__device__ __forceinline__ double px(const T* p, long long i) { return (double)p[i]; }

// dst[n,c,y,x] = bilinear(src[n,c], M_n (x, y, 1)) -- skimage.transform.warp(order=1, mode='reflect') with M the inverse map.
// ROUND: the warped class mask is rounded half-to-even (np.round, augment.py:155) and stored as a uint8 class index.
template <typename TI, typename TO, bool ROUND>
__global__ void __launch_bounds__(TPB) aug_warp_kernel(const TI* __restrict__ src, TO* __restrict__ dst, const double* __restrict__ mats,
                                                       int C, int H, int W) {
  const int plane_id = blockIdx.y;
  const int n = plane_id / C;
  const long long plane = (long long)H * W;
  const TI* s = src + (long long)plane_id * plane;
  TO* d = dst + (long long)plane_id * plane;
  const double m0 = mats[n * 6 + 0], m1 = mats[n * 6 + 1], m2 = mats[n * 6 + 2], m3 = mats[n * 6 + 3], m4 = mats[n * 6 + 4],
               m5 = mats[n * 6 + 5];
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < plane; i += (long long)gridDim.x * TPB) {
    const int y = (int)((unsigned)i / (unsigned)W), x = (int)((unsigned)i - (unsigned)y * (unsigned)W);      // plane < 2^31 (host check)
    const double c = m0 * x + m1 * y + m2;
    const double r = m3 * x + m4 * y + m5;
    const double fr = floor(r), fc = floor(c);
    const long long minr = (long long)fr, minc = (long long)fc, maxr = (long long)ceil(r), maxc = (long long)ceil(c);
    const double dr = r - fr, dc = c - fc;
    // in-range coordinates (almost every sample) skip the 64-bit mirror arithmetic
    const bool inside = minr >= 0 && maxr < H && minc >= 0 && maxc < W;
    const int r0 = inside ? (int)minr : (int)reflect_nodup(H, minr), r1 = inside ? (int)maxr : (int)reflect_nodup(H, maxr);
    const int c0 = inside ? (int)minc : (int)reflect_nodup(W, minc), c1 = inside ? (int)maxc : (int)reflect_nodup(W, maxc);
    const double top = (1.0 - dc) * px(s, r0 * W + c0) + dc * px(s, r0 * W + c1);
    const double bot = (1.0 - dc) * px(s, r1 * W + c0) + dc * px(s, r1 * W + c1);
    const double v = (1.0 - dr) * top + dr * bot;
    if (ROUND) {
      float f = rintf((float)v);
      f = fminf(fmaxf(f, 0.f), 255.f);
      d[i] = (TO)f;
    } else {
      d[i] = (TO)v;
    }
  }
}

// partial[n][b][0 / 1] = min / max over block b's share of image n (all channels)
__global__ void __launch_bounds__(TPB) aug_minmax_kernel(const float* __restrict__ x, float* __restrict__ partial, long long per_image) {
  const float* p = x + (long long)blockIdx.y * per_image;
  float lo = INFINITY, hi = -INFINITY;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < per_image; i += (long long)gridDim.x * TPB) {
    const float v = p[i];
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  __shared__ float slo[TPB / 32], shi[TPB / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    slo[threadIdx.x >> 5] = lo;
    shi[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < TPB / 32; ++w) {
      lo = fminf(lo, slo[w]);
      hi = fmaxf(hi, shi[w]);
    }
    partial[((long long)blockIdx.y * MM_BLOCKS + blockIdx.x) * 2 + 0] = lo;
    partial[((long long)blockIdx.y * MM_BLOCKS + blockIdx.x) * 2 + 1] = hi;
  }
}

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

// x += range_n * (noise_factor_n * z + shift_factor_n), z ~ N(0, 1) (Philox4x32-10 + Box-Muller), range_n = max - min of image n
// (augment.py:118-127 with sigma = noise_factor * range; :141-153 with delta = shift_factor * range)
__global__ void __launch_bounds__(TPB) aug_noise_kernel(float* __restrict__ x, const float* __restrict__ partial,
                                                        const float* __restrict__ factors, long long per_image, unsigned long long seed,
                                                        unsigned long long offset) {
  const int n = blockIdx.y;
  __shared__ float s_range;
  __shared__ float s_lo[MM_BLOCKS / 32], s_hi[MM_BLOCKS / 32];
  if (threadIdx.x < MM_BLOCKS) {          // one partial per thread, shuffle min/max (order-independent)
    float lo = partial[((long long)n * MM_BLOCKS + threadIdx.x) * 2 + 0];
    float hi = partial[((long long)n * MM_BLOCKS + threadIdx.x) * 2 + 1];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
      s_lo[threadIdx.x >> 5] = lo;
      s_hi[threadIdx.x >> 5] = hi;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float lo = s_lo[0], hi = s_hi[0];
    for (int w = 1; w < MM_BLOCKS / 32; ++w) {
      lo = fminf(lo, s_lo[w]);
      hi = fmaxf(hi, s_hi[w]);
    }
    s_range = hi - lo;
  }
  __syncthreads();
  const float sigma = factors[n * 2 + 0] * s_range, delta = factors[n * 2 + 1] * s_range;
  float* p = x + (long long)n * per_image;
  const long long quads = (per_image + 3) / 4;
  for (long long q = (long long)blockIdx.x * TPB + threadIdx.x; q < quads; q += (long long)gridDim.x * TPB) {
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    if (sigma != 0.f) {
      const unsigned long long ctr = offset + (unsigned long long)n * (unsigned long long)quads + (unsigned long long)q;
      uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0x41554721u, 0u};
      philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float u1 = ((float)c[2 * h] + 1.0f) * 2.3283064365386963e-10f;      // (0, 1]
        const float u2 = (float)c[2 * h + 1] * 2.3283064365386963e-10f;
        const float rad = sqrtf(-2.f * __logf(u1));
        float sn, cs;
        __sincosf(6.283185307179586f * u2, &sn, &cs);
        z[2 * h] = rad * cs;
        z[2 * h + 1] = rad * sn;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long i = q * 4 + k;
      if (i < per_image) p[i] = p[i] + sigma * z[k] + delta;
    }
  }
}

// one axis of scipy.ndimage.gaussian_filter(mode='reflect'): dst[i] = sum_k w_n[|k|] src[reflect_dup(i + k)], double accumulate,
// float32 result (correlate1d's output type).  weights: double[N][MAX_RADIUS + 1] (w[0] = centre), radius: int[N]
__global__ void __launch_bounds__(TPB) aug_blur_axis_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                            const double* __restrict__ weights, const int* __restrict__ radius, int axis, int C,
                                                            int H, int W) {
  const int plane_id = blockIdx.y;
  const int n = plane_id / C;
  const long long plane = (long long)H * W;
  const float* s = src + (long long)plane_id * plane;
  float* d = dst + (long long)plane_id * plane;
  const int rad = radius[n];
  const double* w = weights + (long long)n * (MAX_RADIUS + 1);
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < plane; i += (long long)gridDim.x * TPB) {
    if (rad == 0) {
      d[i] = s[i];
      continue;
    }
    const int y = (int)((unsigned)i / (unsigned)W), x = (int)((unsigned)i - (unsigned)y * (unsigned)W);
    double acc = 0.0;
    if (axis == 0) {
      for (int k = -rad; k <= rad; ++k) acc += w[k < 0 ? -k : k] * (double)s[(long long)reflect_dup(H, y + k) * W + x];
    } else {
      for (int k = -rad; k <= rad; ++k) acc += w[k < 0 ? -k : k] * (double)s[(long long)y * W + reflect_dup(W, x + k)];
    }
    d[i] = (float)acc;
  }
}

// the channel axis of the same filter (the reference blurs the H x W x C array along ALL axes, augment.py:139):
// x[n, c, p] <- sum_c' mix[n][c][c'] x[n, c', p]
__global__ void __launch_bounds__(TPB) aug_chanmix_kernel(float* __restrict__ x, const double* __restrict__ mix, int C, long long plane) {
  const int n = blockIdx.y;
  float* p = x + (long long)n * C * plane;
  const double* m = mix + (long long)n * C * C;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < plane; i += (long long)gridDim.x * TPB) {
    double v[UB_MAX_CHANNELS], o[UB_MAX_CHANNELS];
    for (int c = 0; c < C; ++c) v[c] = (double)p[(long long)c * plane + i];
    for (int c = 0; c < C; ++c) {
      o[c] = 0.0;
      for (int k = 0; k < C; ++k) o[c] += m[c * C + k] * v[k];
    }
    for (int c = 0; c < C; ++c) p[(long long)c * plane + i] = (float)o[c];
  }
}

inline int blocks_for(long long items, int cap) {
  long long b = (items + TPB - 1) / TPB;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

extern "C" {

int ub_aug_warp(const void* src, int src_dtype, void* dst, int dst_dtype, const double* mats, int N, int C, int H, int W,
                cudaStream_t stream) {
  UB_CHECK_ARG(src && dst && mats && N > 0 && C > 0 && H > 0 && W > 0, "aug_warp: bad args");
  UB_CHECK_ARG(src != dst, "aug_warp: in-place warps are not possible");
  UB_CHECK_SHAPE((long long)N * C <= 65535, "aug_warp: N*C=%lld planes exceed the grid limit", (long long)N * C);
  UB_CHECK_SHAPE((long long)H * W < (1ll << 31), "aug_warp: planes of %d x %d pixels exceed 2^31", H, W);
  const dim3 grid(blocks_for((long long)H * W, ub_num_sms() * 8), N * C);
  if (dst_dtype == 2) {
    if (src_dtype == 0) aug_warp_kernel<uint8_t, float, false><<<grid, TPB, 0, stream>>>((const uint8_t*)src, (float*)dst, mats, C, H, W);
    else if (src_dtype == 1) aug_warp_kernel<uint16_t, float, false><<<grid, TPB, 0, stream>>>((const uint16_t*)src, (float*)dst, mats, C, H, W);
    else if (src_dtype == 2) aug_warp_kernel<float, float, false><<<grid, TPB, 0, stream>>>((const float*)src, (float*)dst, mats, C, H, W);
    else UB_CHECK_ARG(false, "aug_warp: src_dtype %d (0 = u8, 1 = u16, 2 = f32)", src_dtype);
  } else if (dst_dtype == 0) {
    if (src_dtype == 0) aug_warp_kernel<uint8_t, uint8_t, true><<<grid, TPB, 0, stream>>>((const uint8_t*)src, (uint8_t*)dst, mats, C, H, W);
    else if (src_dtype == 2) aug_warp_kernel<float, uint8_t, true><<<grid, TPB, 0, stream>>>((const float*)src, (uint8_t*)dst, mats, C, H, W);
    else UB_CHECK_ARG(false, "aug_warp: rounding output needs a u8 or f32 source, got %d", src_dtype);
  } else {
    UB_CHECK_ARG(false, "aug_warp: dst_dtype %d (2 = f32, 0 = u8 rounded)", dst_dtype);
  }
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_aug_minmax(const float* x, float* partial, int N, long long per_image, cudaStream_t stream) {
  UB_CHECK_ARG(x && partial && N > 0 && per_image > 0, "aug_minmax: bad args");
  UB_CHECK_SHAPE(N <= 65535, "aug_minmax: N");
  aug_minmax_kernel<<<dim3(MM_BLOCKS, N), TPB, 0, stream>>>(x, partial, per_image);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_aug_noise(float* x, const float* minmax_partial, const float* factors, int N, long long per_image, unsigned long long seed,
                 unsigned long long offset, cudaStream_t stream) {
  UB_CHECK_ARG(x && minmax_partial && factors && N > 0 && per_image > 0, "aug_noise: bad args");
  UB_CHECK_SHAPE(N <= 65535, "aug_noise: N");
  aug_noise_kernel<<<dim3(blocks_for((per_image + 3) / 4, ub_num_sms() * 4), N), TPB, 0, stream>>>(x, minmax_partial, factors, per_image, seed,
                                                                                               offset);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_aug_blur_axis(const float* src, float* dst, const double* weights, const int* radius, int axis, int N, int C, int H, int W,
                     cudaStream_t stream) {
  UB_CHECK_ARG(src && dst && weights && radius && src != dst && (axis == 0 || axis == 1) && N > 0 && C > 0 && H > 0 && W > 0,
               "aug_blur_axis: bad args");
  UB_CHECK_SHAPE((long long)N * C <= 65535 && (long long)H * W < (1ll << 31), "aug_blur_axis: N*C planes / plane size");
  aug_blur_axis_kernel<<<dim3(blocks_for((long long)H * W, ub_num_sms() * 8), N * C), TPB, 0, stream>>>(src, dst, weights, radius, axis, C, H, W);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_aug_chanmix(float* x, const double* mix, int N, int C, long long plane, cudaStream_t stream) {
  UB_CHECK_ARG(x && mix && N > 0 && C >= 1 && C <= UB_MAX_CHANNELS && plane > 0, "aug_chanmix: bad args");
  UB_CHECK_SHAPE(N <= 65535, "aug_chanmix: N");
  aug_chanmix_kernel<<<dim3(blocks_for(plane, ub_num_sms() * 8), N), TPB, 0, stream>>>(x, mix, C, plane);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

}  // extern "C"
