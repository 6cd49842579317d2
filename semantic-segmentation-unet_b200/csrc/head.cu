// First layer (Cin <= UB_MAX_CHANNELS: templated for 1..4 channels, run-time channel loops above; not a tensor-core problem) and the class head of the U-Net (sm_100a, HBM-bound).
//
//   conv_first : relu(conv3x3_same(x_nchw_fp32, W) + b) -> NHWC 64 channels (+BN stat partials)   UNet/model.py:88
//   head_fwd   : relu(conv1x1(x, W) + b), 64 -> K classes, fp32 [P][K] (+BN stat partials)          UNet/model.py:136
//   head_loss  : BN -> softmax -> cross-entropy, argmax, accuracy count, dL/dlogits                 UNet/model.py:139-142, 211-215
//   head_bwd   : BN backward + ReLU mask + 1x1 dgrad / wgrad / bias grad
//   head_argmax: inference epilogue -- BN(moving stats folded) -> argmax written straight into the tile's zone of the
//                output mask (UNet/inference.py:105-129), optional softmax for the model-call contract
// Thread mapping for the 64-channel tensors: 8 threads per pixel x 8 channels each (128-bit bf16 accesses, a warp
// covers 4 pixels = 512 contiguous bytes); class-dimension reductions use warp shuffles.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int TPB = 256;
constexpr int KMAX = UB_MAX_CLASSES;

template <typename T>
__device__ __forceinline__ void ld8(const T* p, float (&f)[8]);
template <>
__device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
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
__device__ __forceinline__ void ld8<float>(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <typename T>
__device__ __forceinline__ void st8(T* p, const float (&f)[8]);
template <>
__device__ __forceinline__ void st8<__nv_bfloat16>(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
template <>
__device__ __forceinline__ void st8<float>(float* p, const float (&f)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}

// sum `v` over all threads of the block; result valid in thread 0
__device__ __forceinline__ float block_sum(float v, float* sh /*[TPB/32]*/) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0)
    for (int w = 0; w < TPB / 32; ++w) t += sh[w];
  return t;
}

// ------------------------------------------------------------------ first conv
// x: fp32 NCHW [N][CIN][H][W]; w: fp32 [64][9][CIN]; out: NHWC [P][64]; partial: [rows][2][64]
template <typename T, int CIN_T>
__global__ void __launch_bounds__(TPB) conv_first_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                             const float* __restrict__ post_scale, const float* __restrict__ post_shift,
                                                             T* __restrict__ out, float* __restrict__ partial, int N, int H, int W,
                                                             int cin_rt) {
  constexpr int CMAX = CIN_T ? CIN_T : UB_MAX_CHANNELS;      // CIN_T == 0: channel count at run time (5 .. UB_MAX_CHANNELS)
  const int CIN = CIN_T ? CIN_T : cin_rt;
  __shared__ float ws[9 * CMAX][64];
  __shared__ float red[TPB * 8];
  for (int i = threadIdx.x; i < 64 * 9 * CIN; i += TPB) {
    const int co = i / (9 * CIN), k = i % (9 * CIN);
    ws[k][co] = w[i];
  }
  __syncthreads();
  const int sub = threadIdx.x & 7, pl = threadIdx.x >> 3;
  float b[8], psc[8], psh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    b[i] = bias[sub * 8 + i];
    psc[i] = post_scale ? post_scale[sub * 8 + i] : 1.f;      // inference: BatchNorm moving statistics folded in
    psh[i] = post_shift ? post_shift[sub * 8 + i] : 0.f;
  }
  float acc_s[8] = {}, acc_q[8] = {};
  const long long P = (long long)N * H * W;
  const long long plane = (long long)H * W;
  for (long long px = (long long)blockIdx.x * 32 + pl; px < P; px += (long long)gridDim.x * 32) {
    const int n = (int)(px / plane);
    const int rem = (int)(px - (long long)n * plane);
    const int h = rem / W, wv = rem % W;
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = b[i];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int hh = h + t / 3 - 1, ww = wv + t % 3 - 1;
      const bool in = (hh >= 0) && (hh < H) && (ww >= 0) && (ww < W);
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        const float xv = in ? __ldg(x + ((long long)n * CIN + ci) * plane + (long long)hh * W + ww) : 0.f;
        const float4 w0 = *reinterpret_cast<const float4*>(&ws[t * CIN + ci][sub * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&ws[t * CIN + ci][sub * 8 + 4]);
        o[0] += xv * w0.x; o[1] += xv * w0.y; o[2] += xv * w0.z; o[3] += xv * w0.w;
        o[4] += xv * w1.x; o[5] += xv * w1.y; o[6] += xv * w1.z; o[7] += xv * w1.w;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i] = fmaxf(o[i], 0.f);
      acc_s[i] += o[i];
      acc_q[i] += o[i] * o[i];
      o[i] = fmaf(o[i], psc[i], psh[i]);
    }
    st8<T>(out + px * 64 + sub * 8, o);
  }
  if (partial) {
#pragma unroll
    for (int comp = 0; comp < 2; ++comp) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 8; ++i) red[pl * 64 + sub * 8 + i] = comp ? acc_q[i] : acc_s[i];
      __syncthreads();
      if (threadIdx.x < 64) {
        float t = 0.f;
        for (int l = 0; l < 32; ++l) t += red[l * 64 + threadIdx.x];
        partial[((size_t)blockIdx.x * 2 + comp) * 64 + threadIdx.x] = t;
      }
    }
  }
}

// Strip-mined variants (W % 4 == 0): one thread = 4 consecutive pixels of a row x 8 output channels.  The 3 x 6 input
// window is loaded once per strip (one 128-bit + two scalar loads per row) and the weights are read from shared memory
// once per strip instead of once per pixel: ~2.5x fewer issued instructions per output than the per-pixel kernels,
// which were issue-bound at 1/8 of the HBM roofline (profiles/r01a).
template <typename T, int CIN_T>
__global__ void __launch_bounds__(TPB) conv_first_fwd_strip_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                   const float* __restrict__ bias, const float* __restrict__ post_scale,
                                                                   const float* __restrict__ post_shift, T* __restrict__ out,
                                                                   float* __restrict__ partial, int N, int H, int W, int cin_rt) {
  constexpr int CMAX = CIN_T ? CIN_T : UB_MAX_CHANNELS;
  const int CIN = CIN_T ? CIN_T : cin_rt;
  __shared__ float ws[9 * CMAX][64];
  __shared__ float red[TPB * 8];
  for (int i = threadIdx.x; i < 64 * 9 * CIN; i += TPB) {
    const int co = i / (9 * CIN), k = i % (9 * CIN);     // w is [co][tap][ci]; k = tap * CIN + ci
    ws[k][co] = w[i];
  }
  __syncthreads();
  const int sub = threadIdx.x & 7, pl = threadIdx.x >> 3;
  float b[8], psc[8], psh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    b[i] = bias[sub * 8 + i];
    psc[i] = post_scale ? post_scale[sub * 8 + i] : 1.f;      // inference: BatchNorm moving statistics folded in
    psh[i] = post_shift ? post_shift[sub * 8 + i] : 0.f;
  }
  float acc_s[8] = {}, acc_q[8] = {};
  const int W4 = W >> 2;
  const long long strips = (long long)N * H * W4;
  const long long plane = (long long)H * W;
  for (long long st = (long long)blockIdx.x * 32 + pl; st < strips; st += (long long)gridDim.x * 32) {
    const int w0 = (int)(st % W4) * 4;
    const long long t = st / W4;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    float o[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int i = 0; i < 8; ++i) o[q][i] = b[i];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      const float* xp = x + ((long long)n * CIN + ci) * plane;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int hh = h + r - 1;
        float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (hh >= 0 && hh < H) {
          const float* row = xp + (long long)hh * W + w0;
          const float4 m = __ldg(reinterpret_cast<const float4*>(row));
          v[1] = m.x; v[2] = m.y; v[3] = m.z; v[4] = m.w;
          if (w0 > 0) v[0] = __ldg(row - 1);
          if (w0 + 4 < W) v[5] = __ldg(row + 4);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int k = (r * 3 + c) * CIN + ci;
          const float4 wa = *reinterpret_cast<const float4*>(&ws[k][sub * 8]);
          const float4 wb = *reinterpret_cast<const float4*>(&ws[k][sub * 8 + 4]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float xv = v[q + c];
            o[q][0] += xv * wa.x; o[q][1] += xv * wa.y; o[q][2] += xv * wa.z; o[q][3] += xv * wa.w;
            o[q][4] += xv * wb.x; o[q][5] += xv * wb.y; o[q][6] += xv * wb.z; o[q][7] += xv * wb.w;
          }
        }
      }
    }
    const long long px0 = ((long long)n * H + h) * W + w0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        o[q][i] = fmaxf(o[q][i], 0.f);
        acc_s[i] += o[q][i];
        acc_q[i] += o[q][i] * o[q][i];
        o[q][i] = fmaf(o[q][i], psc[i], psh[i]);
      }
      st8<T>(out + (px0 + q) * 64 + sub * 8, o[q]);
    }
  }
  if (partial) {
#pragma unroll
    for (int comp = 0; comp < 2; ++comp) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 8; ++i) red[pl * 64 + sub * 8 + i] = comp ? acc_q[i] : acc_s[i];
      __syncthreads();
      if (threadIdx.x < 64) {
        float t = 0.f;
        for (int l = 0; l < 32; ++l) t += red[l * 64 + threadIdx.x];
        partial[((size_t)blockIdx.x * 2 + comp) * 64 + threadIdx.x] = t;
      }
    }
  }
}

// Inference tiles read IN PLACE from the resident, normalised image (UNet/inference.py:46, :97-105): tile n covers image rows
// [origin[2n], +H) and columns [origin[2n+1], +W) of img [CIN][rows >= img_h][pitch]; source coordinates past the unpadded extent
// (img_h, img_w) are mirrored without repeating the edge (np.pad mode='reflect', the bottom/right padding to a multiple of 16), and
// positions outside the TILE are zero (the tile is convolved as an image of its own, 'same' padding).  Same strip mining as above.
__device__ __forceinline__ int reflect_index(int i, int n) { return i < n ? i : 2 * (n - 1) - i; }

template <typename T, int CIN_T>
__global__ void __launch_bounds__(TPB) conv_first_fwd_tiles_kernel(const float* __restrict__ img, const int* __restrict__ origin, int img_h,
                                                                   int img_w, long long pitch, long long plane_stride,
                                                                   const float* __restrict__ w, const float* __restrict__ bias,
                                                                   const float* __restrict__ post_scale, const float* __restrict__ post_shift,
                                                                   T* __restrict__ out, int N, int H, int W, int cin_rt) {
  constexpr int CMAX = CIN_T ? CIN_T : UB_MAX_CHANNELS;
  const int CIN = CIN_T ? CIN_T : cin_rt;
  __shared__ float ws[9 * CMAX][64];
  for (int i = threadIdx.x; i < 64 * 9 * CIN; i += TPB) {
    const int co = i / (9 * CIN), k = i % (9 * CIN);
    ws[k][co] = w[i];
  }
  __syncthreads();
  const int sub = threadIdx.x & 7, pl = threadIdx.x >> 3;
  float b[8], psc[8], psh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    b[i] = bias[sub * 8 + i];
    psc[i] = post_scale[sub * 8 + i];
    psh[i] = post_shift[sub * 8 + i];
  }
  const int W4 = W >> 2;
  const long long strips = (long long)N * H * W4;
  for (long long st = (long long)blockIdx.x * 32 + pl; st < strips; st += (long long)gridDim.x * 32) {
    const int w0 = (int)(st % W4) * 4;
    const long long t = st / W4;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    const int oy = __ldg(origin + 2 * n), ox = __ldg(origin + 2 * n + 1);
    float o[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int i = 0; i < 8; ++i) o[q][i] = b[i];
    const int sx = ox + w0;
    const bool vec = (sx + 4 <= img_w) && (((pitch | plane_stride | sx) & 3) == 0);      // 16-byte aligned rows in every channel plane
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      const float* xp = img + (long long)ci * plane_stride;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int hh = h + r - 1;
        float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (hh >= 0 && hh < H) {
          const float* row = xp + (long long)reflect_index(oy + hh, img_h) * pitch;
          if (vec) {
            const float4 m = __ldg(reinterpret_cast<const float4*>(row + sx));
            v[1] = m.x; v[2] = m.y; v[3] = m.z; v[4] = m.w;
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) v[1 + q] = __ldg(row + reflect_index(sx + q, img_w));
          }
          if (w0 > 0) v[0] = __ldg(row + reflect_index(sx - 1, img_w));
          if (w0 + 4 < W) v[5] = __ldg(row + reflect_index(sx + 4, img_w));
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int k = (r * 3 + c) * CIN + ci;
          const float4 wa = *reinterpret_cast<const float4*>(&ws[k][sub * 8]);
          const float4 wb = *reinterpret_cast<const float4*>(&ws[k][sub * 8 + 4]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float xv = v[q + c];
            o[q][0] += xv * wa.x; o[q][1] += xv * wa.y; o[q][2] += xv * wa.z; o[q][3] += xv * wa.w;
            o[q][4] += xv * wb.x; o[q][5] += xv * wb.y; o[q][6] += xv * wb.z; o[q][7] += xv * wb.w;
          }
        }
      }
    }
    const long long px0 = ((long long)n * H + h) * W + w0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[q][i] = fmaf(fmaxf(o[q][i], 0.f), psc[i], psh[i]);
      st8<T>(out + (px0 + q) * 64 + sub * 8, o[q]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(TPB, 2) conv_first_wgrad_strip_kernel(const float* __restrict__ x, const T* __restrict__ dz,
                                                                     float* __restrict__ partial, int N, int H, int W, int CIN) {
  __shared__ float red[TPB * 8];
  const int ci = blockIdx.y;
  const int sub = threadIdx.x & 7, pl = threadIdx.x >> 3;
  float acc[9][8] = {};
  const int W4 = W >> 2;
  const long long strips = (long long)N * H * W4;
  const long long plane = (long long)H * W;
  for (long long st = (long long)blockIdx.x * 32 + pl; st < strips; st += (long long)gridDim.x * 32) {
    const int w0 = (int)(st % W4) * 4;
    const long long t = st / W4;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    const long long px0 = ((long long)n * H + h) * W + w0;
    float d[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q) ld8<T>(dz + (px0 + q) * 64 + sub * 8, d[q]);
    const float* xp = x + ((long long)n * CIN + ci) * plane;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int hh = h + r - 1;
      float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (hh >= 0 && hh < H) {
        const float* row = xp + (long long)hh * W + w0;
        const float4 m = __ldg(reinterpret_cast<const float4*>(row));
        v[1] = m.x; v[2] = m.y; v[3] = m.z; v[4] = m.w;
        if (w0 > 0) v[0] = __ldg(row - 1);
        if (w0 + 4 < W) v[5] = __ldg(row + 4);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float xv = v[q + c];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[r * 3 + c][i] += xv * d[q][i];
        }
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) red[pl * 64 + sub * 8 + i] = acc[t][i];
    __syncthreads();
    if (threadIdx.x < 64) {
      float s = 0.f;
      for (int l = 0; l < 32; ++l) s += red[l * 64 + threadIdx.x];
      partial[(((size_t)blockIdx.x * CIN + ci) * 9 + t) * 64 + threadIdx.x] = s;
    }
  }
}

// dx[n][ci][h][w] = sum_{tap, co} dz[(h, w) - shift(tap)][co] * w[co][tap][ci]   (gradient w.r.t. the network input; only
// UNet.estimate_radius needs it -- training never differentiates the image).  One thread per input pixel and channel.
template <typename T>
__global__ void __launch_bounds__(TPB) conv_first_dgrad_kernel(const T* __restrict__ dz, const float* __restrict__ w, float* __restrict__ dx,
                                                               int N, int H, int W, int CIN) {
  const long long total = (long long)N * CIN * H * W;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < total; i += (long long)gridDim.x * TPB) {
    const int wv = (int)(i % W);
    long long t = i / W;
    const int h = (int)(t % H); t /= H;
    const int ci = (int)(t % CIN);
    const int n = (int)(t / CIN);
    float acc = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
      const int hh = h - (tap / 3 - 1), ww = wv - (tap % 3 - 1);       // output pixel that read this input through `tap`
      if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
      const T* d = dz + (((long long)n * H + hh) * W + ww) * 64;
      for (int c8 = 0; c8 < 8; ++c8) {
        float f[8];
        ld8<T>(d + c8 * 8, f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += f[k] * w[((c8 * 8 + k) * 9 + tap) * CIN + ci];
      }
    }
    dx[i] = acc;
  }
}

// dW partial[row][ci][tap][64] = sum_p dz[p][co] * x[p + tap][ci]   (blockIdx.y = ci)
template <typename T>
__global__ void __launch_bounds__(TPB) conv_first_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ dz, float* __restrict__ partial,
                                                               int N, int H, int W, int CIN) {
  __shared__ float red[TPB * 8];
  const int ci = blockIdx.y;
  const int sub = threadIdx.x & 7, pl = threadIdx.x >> 3;
  float acc[9][8] = {};
  const long long P = (long long)N * H * W;
  const long long plane = (long long)H * W;
  for (long long px = (long long)blockIdx.x * 32 + pl; px < P; px += (long long)gridDim.x * 32) {
    const int n = (int)(px / plane);
    const int rem = (int)(px - (long long)n * plane);
    const int h = rem / W, wv = rem % W;
    float d[8];
    ld8<T>(dz + px * 64 + sub * 8, d);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int hh = h + t / 3 - 1, ww = wv + t % 3 - 1;
      const bool in = (hh >= 0) && (hh < H) && (ww >= 0) && (ww < W);
      const float xv = in ? __ldg(x + ((long long)n * CIN + ci) * plane + (long long)hh * W + ww) : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[t][i] += xv * d[i];
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) red[pl * 64 + sub * 8 + i] = acc[t][i];
    __syncthreads();
    if (threadIdx.x < 64) {
      float s = 0.f;
      for (int l = 0; l < 32; ++l) s += red[l * 64 + threadIdx.x];
      partial[(((size_t)blockIdx.x * CIN + ci) * 9 + t) * 64 + threadIdx.x] = s;
    }
  }
}
// dw[co][tap][ci] = sum_rows partial[row][ci][tap][co]
__global__ void conv_first_wgrad_finalize_kernel(const float* __restrict__ partial, float* __restrict__ dw, int rows, int CIN) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 9 * CIN) return;
  const int co = i / (9 * CIN), r = i % (9 * CIN), t = r / CIN, ci = r % CIN;
  double s = 0.0;
  for (int row = 0; row < rows; ++row) s += (double)partial[(((size_t)row * CIN + ci) * 9 + t) * 64 + co];
  dw[i] = (float)s;
}

// ------------------------------------------------------------------ head forward: a[P][K] = relu(x[P][64] . w[K][64] + b)
// K is a template parameter (weights stay in registers); two pixels per thread per iteration keep two 16-byte loads in flight.
template <typename T, int K>
__global__ void __launch_bounds__(TPB) head_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                       float* __restrict__ a_out, float* __restrict__ partial, long long P) {
  __shared__ float sh[TPB / 32];
  const int sub = threadIdx.x & 7, pl = threadIdx.x >> 3;
  float wr[K][8], bk[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    bk[k] = b[k];
#pragma unroll
    for (int i = 0; i < 8; ++i) wr[k][i] = w[k * 64 + sub * 8 + i];
  }
  float s[K] = {}, q[K] = {};
  const long long iters = (P + 63) / 64;
  for (long long it = blockIdx.x; it < iters; it += gridDim.x) {
    long long px[2];
    bool ok[2];
    float f[2][8] = {};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      px[u] = it * 64 + u * 32 + pl;
      ok[u] = px[u] < P;
      if (ok[u]) ld8<T>(x + px[u] * 64 + sub * 8, f[u]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) d += wr[k][i] * f[u][i];
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        const float a = fmaxf(d + bk[k], 0.f);
        if (ok[u] && sub == (k & 7)) {
          a_out[px[u] * K + k] = a;
          s[k] += a;
          q[k] += a * a;
        }
      }
    }
  }
  if (partial) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float ts = block_sum(s[k], sh);
      const float tq = block_sum(q[k], sh);
      if (threadIdx.x == 0) {
        partial[((size_t)blockIdx.x * 2 + 0) * K + k] = ts;
        partial[((size_t)blockIdx.x * 2 + 1) * K + k] = tq;
      }
    }
  }
}

// ------------------------------------------------------------------ head loss
// per pixel: y = BN(a); p = softmax(y); ce = -log p[label]; partial[row] = {inv_denom * sum ce * cw[label], acc_scale * #correct}
// dlogits = (p - onehot) * cw[label] * inv_denom ; optional softmax output
__global__ void __launch_bounds__(TPB) head_loss_kernel(const float* __restrict__ a, const float* __restrict__ mean, const float* __restrict__ rstd,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        const uint8_t* __restrict__ labels, const float* __restrict__ class_w, float inv_denom,
                                                        float acc_scale, float smooth, float* __restrict__ softmax_out, float* __restrict__ dlogits,
                                                        float* __restrict__ partial, long long P, int K) {
  __shared__ float sh[TPB / 32];
  float sc[KMAX], sf[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    sc[k] = (k < K) ? gamma[k] * rstd[k] : 0.f;
    sf[k] = (k < K) ? beta[k] - mean[k] * sc[k] : 0.f;
  }
  float loss = 0.f, correct = 0.f;
  for (long long px = (long long)blockIdx.x * TPB + threadIdx.x; px < P; px += (long long)gridDim.x * TPB) {
    float y[KMAX];
    float mx = -INFINITY;
    int am = 0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < K) {
        y[k] = a[px * K + k] * sc[k] + sf[k];
        if (y[k] > mx) {
          mx = y[k];
          am = k;
        }
      }
    }
    // label smoothing (Keras CategoricalCrossentropy, UNet/model.py:77): target t'_k = (1 - eps) [k == label] + eps / K, so the loss also
    // needs sum_k log p_k = sum_k (y_k - max) - K log(sum exp)
    float se = 0.f, sy = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) {
        sy += y[k] - mx;
        y[k] = __expf(y[k] - mx);
        se += y[k];
      }
    const float inv = 1.f / se;
    const int lab = labels ? (int)labels[px] : 0;
    const float cw = class_w ? class_w[lab] : 1.f;
    const float t_on = 1.f - smooth + smooth / (float)K, t_off = smooth / (float)K;
    float pl = 1.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) {
        const float p = y[k] * inv;
        if (k == lab) pl = p;
        if (softmax_out) softmax_out[px * K + k] = p;
        if (dlogits) dlogits[px * K + k] = (p - (k == lab ? t_on : t_off)) * cw * inv_denom;
      }
    if (labels) {
      float l = -__logf(fmaxf(pl, 1e-37f)) * (1.f - smooth);
      if (smooth != 0.f) l -= t_off * (sy - (float)K * __logf(se));
      loss += l * cw;
      correct += (am == lab) ? 1.f : 0.f;
    }
  }
  if (partial) {
    const float tl = block_sum(loss, sh);
    const float tc = block_sum(correct, sh);
    if (threadIdx.x == 0) {
      partial[(size_t)blockIdx.x * 2 + 0] = tl * inv_denom;
      partial[(size_t)blockIdx.x * 2 + 1] = tc * acc_scale;
    }
  }
}

// one-hot int32 [P][K] -> uint8 class index (argmax, first max): the reference's label contract (imagereader.py:302-312)
__global__ void __launch_bounds__(TPB) onehot_to_index_kernel(const int* __restrict__ onehot, uint8_t* __restrict__ idx, long long P, int K) {
  for (long long px = (long long)blockIdx.x * TPB + threadIdx.x; px < P; px += (long long)gridDim.x * TPB) {
    int best = onehot[px * K], am = 0;
    for (int k = 1; k < K; ++k) {
      const int v = onehot[px * K + k];
      if (v > best) {
        best = v;
        am = k;
      }
    }
    idx[px] = (uint8_t)am;
  }
}

// ------------------------------------------------------------------ head backward
// pass 1: partial[row][0][k] = sum dy_k ; partial[row][1][k] = sum dy_k * xhat_k
__global__ void __launch_bounds__(TPB) head_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ a, const float* __restrict__ mean,
                                                              const float* __restrict__ rstd, float* __restrict__ partial, long long P, int K) {
  __shared__ float sh[TPB / 32];
  float s[KMAX] = {}, q[KMAX] = {};
  for (long long px = (long long)blockIdx.x * TPB + threadIdx.x; px < P; px += (long long)gridDim.x * TPB) {
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) {
        const float d = dy[px * K + k];
        s[k] += d;
        q[k] += d * (a[px * K + k] - mean[k]) * rstd[k];
      }
  }
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    if (k < K) {
      const float ts = block_sum(s[k], sh);
      const float tq = block_sum(q[k], sh);
      if (threadIdx.x == 0) {
        partial[((size_t)blockIdx.x * 2 + 0) * K + k] = ts;
        partial[((size_t)blockIdx.x * 2 + 1) * K + k] = tq;
      }
    }
}

// pass 2: dz_k = gamma_k rstd_k (dy_k - dbeta_k/P - xhat_k dgamma_k/P) [a_k > 0]
//         dx[p][c] = sum_k w[k][c] dz_k   ;   partial[row] = { dW[k][c] (K*64), db[k] (K) }
template <typename T, int K, bool RED, int PPI = 2>
__global__ void __launch_bounds__(TPB) head_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ a, const T* __restrict__ x,
                                                             const float* __restrict__ w, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                             const float* __restrict__ dbeta, const float* __restrict__ dgamma,
                                                             T* __restrict__ dx, float* __restrict__ partial, long long P,
                                                             const T* __restrict__ red_a, const float* __restrict__ red_mean,
                                                             const float* __restrict__ red_rstd, float* __restrict__ red_out) {
  // RED: dx is dL/dy of the BatchNorm'd 64-channel tensor below the head (dec1b); its backward sums
  // red_out[row][0][c] = sum dx, red_out[row][1][c] = rstd_c * sum dx * (red_a - mean_c) (of the STORED dx) are accumulated here
  __shared__ float red[TPB / 32][K * 64 + K];
  __shared__ float rred[RED ? TPB / 32 : 1][128];
  float racc[2][8] = {};
  float rmu[8] = {};
  if (RED) {
#pragma unroll
    for (int i = 0; i < 8; ++i) rmu[i] = red_mean[(threadIdx.x & 7) * 8 + i];
  }
  const int sub = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float wr[K][8], accw[K][8] = {}, accb[K] = {};
  float g_rs[K], mu[K], rs[K], db[K], dg[K];
  const float invP = 1.f / (float)P;
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int i = 0; i < 8; ++i) wr[k][i] = w[k * 64 + sub * 8 + i];
    mu[k] = mean[k];
    rs[k] = rstd[k];
    g_rs[k] = gamma[k] * rs[k];
    db[k] = dbeta[k] * invP;
    dg[k] = dgamma[k] * invP;
  }
  // PPI pixels per thread per iteration: PPI 16-byte loads of x in flight (94 registers allow 2 blocks per SM: with PPI = 2 the kernel
  // was latency-bound at 3.1 TB/s, profiles/r01_kernels.md)
  const long long iters = (P + 32 * PPI - 1) / (32 * PPI);
  for (long long it = blockIdx.x; it < iters; it += gridDim.x) {
    long long px[PPI];
    bool ok[PPI];
    float f[PPI][8] = {};
    float av[PPI][K], dv[PPI][K];
#pragma unroll
    for (int u = 0; u < PPI; ++u) {
      px[u] = it * (32 * PPI) + u * 32 + pl;
      ok[u] = px[u] < P;
      if (ok[u]) {
        ld8<T>(x + px[u] * 64 + sub * 8, f[u]);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          av[u][k] = __ldg(a + px[u] * K + k);
          dv[u][k] = __ldg(dy + px[u] * K + k);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < PPI; ++u) {
      if (ok[u]) {
        float o[8] = {};
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const float xh = (av[u][k] - mu[k]) * rs[k];
          float dz = g_rs[k] * (dv[u][k] - db[k] - xh * dg[k]);
          if (!(av[u][k] > 0.f)) dz = 0.f;
          if (sub == 0) accb[k] += dz;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            o[i] += wr[k][i] * dz;
            accw[k][i] += dz * f[u][i];
          }
        }
        if (dx) st8<T>(dx + px[u] * 64 + sub * 8, o);
        if (RED) {
          float av[8];
          ld8<T>(red_a + px[u] * 64 + sub * 8, av);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float vs = sizeof(T) == 2 ? __bfloat162float(__float2bfloat16(o[i])) : o[i];
            racc[0][i] += vs;
            racc[1][i] = fmaf(vs, av[i] - rmu[i], racc[1][i]);
          }
        }
      }
    }
  }
  if (RED) {
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v = racc[c][i];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (lane < 8) rred[warp][c * 64 + sub * 8 + i] = v;
      }
  }
  // reduce over the 4 pixels of each warp (lanes with equal sub), then over warps through smem
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = accw[k][i];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 8) red[warp][k * 64 + sub * 8 + i] = v;
    }
    float vb = accb[k];
    vb = warp_sum(vb);
    if (lane == 0) red[warp][K * 64 + k] = vb;
  }
  __syncthreads();
  constexpr int ncomp = K * 64 + K;
  for (int i = threadIdx.x; i < ncomp; i += TPB) {
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < TPB / 32; ++wv) t += red[wv][i];
    partial[(size_t)blockIdx.x * ncomp + i] = t;
  }
  if (RED) {
    for (int i = threadIdx.x; i < 128; i += TPB) {
      float t = 0.f;
#pragma unroll
      for (int wv = 0; wv < TPB / 32; ++wv) t += rred[wv][i];
      if (i >= 64) t *= red_rstd[i - 64];
      red_out[(size_t)blockIdx.x * 128 + i] = t;
    }
  }
}

// ------------------------------------------------------------------ inference epilogue
// logits_k = relu(x . w_k + b_k) * scale_k + shift_k (BN moving stats folded); argmax (first max) -> mask zone.
// A batch of `ntiles` equal-sized h x w tiles (x = [ntiles][h][w][64]); tile t writes its crop box
// [cy0,cy1) x [cx0,cx1) to mask[(dst_y + ty - cy0) * mask_ld + dst_x + tx - cx0]  (geo[t] = {cy0,cy1,cx0,cx1,dst_y,dst_x}).
struct TileGeo {
  int cy0, cy1, cx0, cx1, dst_y, dst_x;
};
template <typename T, int K>
__global__ void __launch_bounds__(TPB) head_argmax_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                          const float* __restrict__ scale, const float* __restrict__ shift, int h, int wd,
                                                          const TileGeo* __restrict__ geo, uint8_t* __restrict__ mask, long long mask_ld,
                                                          float* __restrict__ softmax_out) {
  const int sub = threadIdx.x & 7, pl = threadIdx.x >> 3;
  float wr[K][8], bk[K], sc[K], sf[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    bk[k] = b[k];
    sc[k] = scale[k];
    sf[k] = shift[k];
#pragma unroll
    for (int i = 0; i < 8; ++i) wr[k][i] = w[k * 64 + sub * 8 + i];
  }
  const int tile = blockIdx.y;
  const TileGeo g = geo ? geo[tile] : TileGeo{0, h, 0, wd, 0, 0};
  const long long P = (long long)h * wd;
  const T* xt = x + (long long)tile * P * 64;
  const long long iters = (P + 31) / 32;
  for (long long it = blockIdx.x; it < iters; it += gridDim.x) {
    const long long px = it * 32 + pl;
    const bool ok = px < P;
    float f[8] = {};
    if (ok) ld8<T>(xt + px * 64 + sub * 8, f);
    float y[K];
    float mx = -INFINITY;
    int am = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) d += wr[k][i] * f[i];
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      y[k] = fmaxf(d + bk[k], 0.f) * sc[k] + sf[k];
      if (y[k] > mx) {
        mx = y[k];
        am = k;
      }
    }
    if (ok && sub == 0) {
      const int ty = (int)(px / wd), tx = (int)(px % wd);
      if (mask && ty >= g.cy0 && ty < g.cy1 && tx >= g.cx0 && tx < g.cx1)
        mask[(long long)(g.dst_y + ty - g.cy0) * mask_ld + g.dst_x + tx - g.cx0] = (uint8_t)am;
      if (softmax_out) {
        float se = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          y[k] = __expf(y[k] - mx);
          se += y[k];
        }
        const float inv = 1.f / se;
#pragma unroll
        for (int k = 0; k < K; ++k) softmax_out[((long long)tile * P + px) * K + k] = y[k] * inv;
      }
    }
  }
}

// plain 2x2 max-pool (inference: the BatchNorm is already folded into the producer), 8 channels per thread
template <typename T>
__global__ void __launch_bounds__(TPB) maxpool_fwd_kernel(const T* __restrict__ y, T* __restrict__ pooled, int N, int H, int W, int C) {
  const int G = C >> 3, Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)N * Ho * Wo * G;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < total; i += (long long)gridDim.x * TPB) {
    const int c0 = (int)(i % G) * 8;
    long long t = i / G;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float best[8], f[4][8];
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4)
      ld8<T>(y + (((long long)n * H + 2 * ho + (s4 >> 1)) * W + 2 * wo + (s4 & 1)) * C + c0, f[s4]);
#pragma unroll
    for (int k = 0; k < 8; ++k) best[k] = fmaxf(fmaxf(f[0][k], f[1][k]), fmaxf(f[2][k], f[3][k]));
    st8<T>(pooled + (((long long)n * Ho + ho) * Wo + wo) * C + c0, best);
  }
}

inline int grid_for(long long work_items, int per_block, int cap) {
  long long g = (work_items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

}  // namespace

// class-per-lane kernels for UB_MAX_CLASSES < K <= UB_MAX_CLASSES_ANY (csrc/head_generic.cu)
int ubg_head_fwd(const void* x, const float* w, const float* b, float* a_out, float* partial, long long P, int K, int dtype, cudaStream_t stream);
int ubg_head_argmax(const void* x, const float* w, const float* b, const float* scale, const float* shift, int K, int ntiles, int h, int wd,
                    const int* geo, unsigned char* mask, long long mask_ld, float* softmax_out, int dtype, cudaStream_t stream);
int ubg_head_loss(const float* a, const float* mean, const float* rstd, const float* gamma, const float* beta, const unsigned char* labels,
                  const float* class_w, float inv_denom, float acc_scale, float smooth, float* softmax_out, float* dlogits, float* partial, long long P,
                  int K, cudaStream_t stream);
int ubg_head_bwd_reduce(const float* dy, const float* a, const float* mean, const float* rstd, float* partial, long long P, int K,
                        cudaStream_t stream);
int ubg_head_bwd_apply(const float* dy, const float* a, const void* x, const float* w, const float* mean, const float* rstd, const float* gamma,
                       const float* dbeta, const float* dgamma, void* dx, float* partial, long long P, int K, int dtype, cudaStream_t stream);

#define UB_DISPATCH_T(dtype, ...)                                   \
  do {                                                              \
    if ((dtype) == UB_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else if ((dtype) == UB_F32) { using T = float; __VA_ARGS__; }   \
    else { ub_set_error("bad dtype %d", (int)(dtype)); return UB_ERR_INVALID_ARG; } \
  } while (0)

#define UB_DISPATCH_K(kk, ...)                                      \
  do {                                                              \
    switch (kk) {                                                   \
      case 1: { constexpr int KK = 1; __VA_ARGS__; } break;         \
      case 2: { constexpr int KK = 2; __VA_ARGS__; } break;         \
      case 3: { constexpr int KK = 3; __VA_ARGS__; } break;         \
      case 4: { constexpr int KK = 4; __VA_ARGS__; } break;         \
      case 5: { constexpr int KK = 5; __VA_ARGS__; } break;         \
      case 6: { constexpr int KK = 6; __VA_ARGS__; } break;         \
      case 7: { constexpr int KK = 7; __VA_ARGS__; } break;         \
      default: { constexpr int KK = 8; __VA_ARGS__; } break;        \
    }                                                               \
  } while (0)

#define UB_DISPATCH_CIN(cin, ...)                                   \
  do {                                                              \
    if ((cin) == 1) { constexpr int CIN = 1; __VA_ARGS__; }         \
    else if ((cin) == 2) { constexpr int CIN = 2; __VA_ARGS__; }    \
    else if ((cin) == 3) { constexpr int CIN = 3; __VA_ARGS__; }    \
    else if ((cin) == 4) { constexpr int CIN = 4; __VA_ARGS__; }    \
    else { constexpr int CIN = 0; __VA_ARGS__; }  /* 5 .. UB_MAX_CHANNELS: channel loops at run time */ \
  } while (0)

extern "C" {

static int conv_first_impl(const float* x_nchw, const float* w, const float* bias, const float* post_scale, const float* post_shift, void* out,
                           float* partial, int N, int H, int W, int Cin, int dtype, cudaStream_t stream);

int ub_conv_first_fwd(const float* x_nchw, const float* w, const float* bias, void* out, float* partial, int N, int H, int W, int Cin,
                      int dtype, cudaStream_t stream) {
  return conv_first_impl(x_nchw, w, bias, nullptr, nullptr, out, partial, N, H, W, Cin, dtype, stream);
}

int ub_conv_first_fwd_affine(const float* x_nchw, const float* w, const float* bias, const float* scale, const float* shift, void* out,
                             int N, int H, int W, int Cin, int dtype, cudaStream_t stream) {
  UB_CHECK_ARG(scale && shift, "conv_first_fwd_affine: null pointer");
  return conv_first_impl(x_nchw, w, bias, scale, shift, out, nullptr, N, H, W, Cin, dtype, stream);
}

int ub_conv_first_fwd_affine_tiles(const float* img, const int* origin_yx, int img_h, int img_w, long long row_pitch, long long plane_stride,
                                   const float* w, const float* bias, const float* scale, const float* shift, void* out, int N, int H, int W,
                                   int Cin, int dtype, cudaStream_t stream) {
  UB_CHECK_ARG(img && origin_yx && w && bias && scale && shift && out, "conv_first_fwd_affine_tiles: null pointer");
  UB_CHECK_SHAPE(Cin >= 1 && Cin <= UB_MAX_CHANNELS, "conv_first_fwd_affine_tiles: Cin=%d must be in [1, %d]", Cin, UB_MAX_CHANNELS);
  UB_CHECK_SHAPE(N > 0 && H > 0 && W > 0 && W % 4 == 0, "conv_first_fwd_affine_tiles: tile width must be a multiple of 4 (got %d x %d)", H, W);
  UB_CHECK_SHAPE(img_h >= 2 && img_w >= 2 && row_pitch >= img_w, "conv_first_fwd_affine_tiles: bad image extent %d x %d, pitch %lld", img_h, img_w,
                 row_pitch);
  UB_CHECK_ARG((reinterpret_cast<uintptr_t>(img) & 15) == 0, "conv_first_fwd_affine_tiles: image base must be 16-byte aligned");
  UB_CHECK_ARG(Cin == 1 || plane_stride >= (long long)img_h * row_pitch, "conv_first_fwd_affine_tiles: channel planes overlap");
  const long long P = (long long)N * H * W;
  const int grid = grid_for(P / 4, 32 * 4, UB_STATS_ROWS);
  UB_DISPATCH_T(dtype, UB_DISPATCH_CIN(Cin, (conv_first_fwd_tiles_kernel<T, CIN><<<grid, TPB, 0, stream>>>(img, origin_yx, img_h, img_w, row_pitch, plane_stride,
                                                                                                        w, bias, scale, shift, (T*)out, N, H, W, Cin))));
  UB_LAUNCH_CHECK();
  return UB_OK;
}

static int conv_first_impl(const float* x_nchw, const float* w, const float* bias, const float* post_scale, const float* post_shift, void* out,
                           float* partial, int N, int H, int W, int Cin, int dtype, cudaStream_t stream) {
  UB_CHECK_ARG(x_nchw && w && bias && out, "conv_first_fwd: null pointer");
  UB_CHECK_SHAPE(Cin >= 1 && Cin <= UB_MAX_CHANNELS, "conv_first_fwd: Cin=%d must be in [1, %d]", Cin, UB_MAX_CHANNELS);
  const long long P = (long long)N * H * W;
  const bool strip = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(x_nchw) & 15) == 0);
  const int grid = grid_for(strip ? P / 4 : P, 32 * 4, UB_STATS_ROWS);
  if (partial) UB_CUDA(cudaMemsetAsync(partial, 0, sizeof(float) * UB_STATS_ROWS * 2 * 64, stream));
  if (strip)
    UB_DISPATCH_T(dtype, UB_DISPATCH_CIN(Cin, (conv_first_fwd_strip_kernel<T, CIN><<<grid, TPB, 0, stream>>>(x_nchw, w, bias, post_scale, post_shift, (T*)out, partial, N, H, W, Cin))));
  else
    UB_DISPATCH_T(dtype, UB_DISPATCH_CIN(Cin, (conv_first_fwd_kernel<T, CIN><<<grid, TPB, 0, stream>>>(x_nchw, w, bias, post_scale, post_shift, (T*)out, partial, N, H, W, Cin))));
  UB_LAUNCH_CHECK();
  return UB_OK;
}

// partial scratch: UB_STATS_ROWS * Cin * 9 * 64 floats
int ub_conv_first_wgrad(const float* x_nchw, const void* dz, float* dw, float* partial, int N, int H, int W, int Cin, int dtype,
                        cudaStream_t stream) {
  UB_CHECK_ARG(x_nchw && dz && dw && partial, "conv_first_wgrad: null pointer");
  UB_CHECK_SHAPE(Cin >= 1 && Cin <= UB_MAX_CHANNELS, "conv_first_wgrad: Cin=%d must be in [1, %d]", Cin, UB_MAX_CHANNELS);
  const long long P = (long long)N * H * W;
  const bool strip = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(x_nchw) & 15) == 0);
  const int rows = grid_for(strip ? P / 4 : P, 32 * 4, UB_STATS_ROWS);
  dim3 grid(rows, Cin);
  if (strip)
    UB_DISPATCH_T(dtype, (conv_first_wgrad_strip_kernel<T><<<grid, TPB, 0, stream>>>(x_nchw, (const T*)dz, partial, N, H, W, Cin)));
  else
    UB_DISPATCH_T(dtype, (conv_first_wgrad_kernel<T><<<grid, TPB, 0, stream>>>(x_nchw, (const T*)dz, partial, N, H, W, Cin)));
  UB_LAUNCH_CHECK();
  conv_first_wgrad_finalize_kernel<<<(64 * 9 * Cin + 127) / 128, 128, 0, stream>>>(partial, dw, rows, Cin);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_conv_first_dgrad(const void* dz, const float* w, float* dx_nchw, int N, int H, int W, int Cin, int dtype, cudaStream_t stream) {
  UB_CHECK_ARG(dz && w && dx_nchw, "conv_first_dgrad: null pointer");
  UB_CHECK_SHAPE(Cin >= 1 && Cin <= UB_MAX_CHANNELS && N > 0 && H > 0 && W > 0, "conv_first_dgrad: bad shape");
  const long long total = (long long)N * Cin * H * W;
  UB_DISPATCH_T(dtype, (conv_first_dgrad_kernel<T><<<grid_for(total, TPB, ub_num_sms() * 8), TPB, 0, stream>>>((const T*)dz, w, dx_nchw, N, H, W, Cin)));
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_head_fwd(const void* x, const float* w, const float* b, float* a_out, float* partial, long long P, int K, int dtype,
                cudaStream_t stream) {
  UB_CHECK_ARG(x && w && b && a_out && P > 0, "head_fwd: bad args");
  UB_CHECK_SHAPE(K >= 1 && K <= UB_MAX_CLASSES_ANY, "head_fwd: number_classes=%d must be in [1, %d]", K, UB_MAX_CLASSES_ANY);
  UB_CHECK_ARG(dtype == UB_BF16 || dtype == UB_F32, "head_fwd: bad dtype %d", dtype);
  if (K > KMAX) return ubg_head_fwd(x, w, b, a_out, partial, P, K, dtype, stream);
  const int grid = grid_for(P, 64 * 4, UB_STATS_ROWS);
  if (partial) UB_CUDA(cudaMemsetAsync(partial, 0, sizeof(float) * UB_STATS_ROWS * 2 * K, stream));
  UB_DISPATCH_T(dtype, UB_DISPATCH_K(K, (head_fwd_kernel<T, KK><<<grid, TPB, 0, stream>>>((const T*)x, w, b, a_out, partial, P))));
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_head_loss(const float* a, const float* mean, const float* rstd, const float* gamma, const float* beta, const unsigned char* labels,
                 const float* class_w, float inv_denom, float acc_scale, float label_smoothing, float* softmax_out, float* dlogits, float* partial,
                 long long P, int K, cudaStream_t stream) {
  UB_CHECK_ARG(a && mean && rstd && gamma && beta && P > 0, "head_loss: bad args");
  UB_CHECK_ARG(label_smoothing >= 0.f && label_smoothing <= 1.f, "head_loss: label_smoothing must be in [0, 1]");
  UB_CHECK_SHAPE(K >= 1 && K <= UB_MAX_CLASSES_ANY, "head_loss: number_classes=%d must be in [1, %d]", K, UB_MAX_CLASSES_ANY);
  if (K > KMAX)
    return ubg_head_loss(a, mean, rstd, gamma, beta, labels, class_w, inv_denom, acc_scale, label_smoothing, softmax_out, dlogits, partial, P, K, stream);
  const int grid = grid_for(P, TPB * 4, UB_STATS_ROWS);
  if (partial) UB_CUDA(cudaMemsetAsync(partial, 0, sizeof(float) * UB_STATS_ROWS * 2, stream));
  head_loss_kernel<<<grid, TPB, 0, stream>>>(a, mean, rstd, gamma, beta, labels, class_w, inv_denom, acc_scale, label_smoothing, softmax_out, dlogits,
                                             partial, P, K);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_onehot_to_index(const int* onehot, unsigned char* idx, long long P, int K, cudaStream_t stream) {
  UB_CHECK_ARG(onehot && idx && P > 0 && K >= 1 && K <= 255, "onehot_to_index: bad args");
  onehot_to_index_kernel<<<grid_for(P, TPB, ub_num_sms() * 8), TPB, 0, stream>>>(onehot, idx, P, K);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_head_bwd_reduce(const float* dy, const float* a, const float* mean, const float* rstd, float* partial, long long P, int K,
                       cudaStream_t stream) {
  UB_CHECK_ARG(dy && a && mean && rstd && partial && P > 0, "head_bwd_reduce: bad args");
  UB_CHECK_SHAPE(K >= 1 && K <= UB_MAX_CLASSES_ANY, "head_bwd_reduce: K");
  if (K > KMAX) return ubg_head_bwd_reduce(dy, a, mean, rstd, partial, P, K, stream);
  const int grid = grid_for(P, TPB * 4, UB_STATS_ROWS);
  UB_CUDA(cudaMemsetAsync(partial, 0, sizeof(float) * UB_STATS_ROWS * 2 * K, stream));
  head_bwd_reduce_kernel<<<grid, TPB, 0, stream>>>(dy, a, mean, rstd, partial, P, K);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

// partial: UB_STATS_ROWS * (K*64 + K) floats; row layout {dW[K][64], db[K]}
static int head_bwd_apply_launch(const float* dy, const float* a, const void* x, const float* w, const float* mean, const float* rstd,
                                 const float* gamma, const float* dbeta, const float* dgamma, void* dx, float* partial, long long P, int K,
                                 int dtype, const void* red_a, const float* red_mean, const float* red_rstd, float* red_out,
                                 cudaStream_t stream) {
  UB_CHECK_ARG(dy && a && x && w && mean && rstd && gamma && dbeta && dgamma && partial && P > 0, "head_bwd_apply: bad args");
  UB_CHECK_SHAPE(K >= 1 && K <= UB_MAX_CLASSES_ANY, "head_bwd_apply: K");
  UB_CHECK_ARG(dtype == UB_BF16 || dtype == UB_F32, "head_bwd_apply: bad dtype %d", dtype);
  if (K > KMAX) {
    UB_CHECK_SHAPE(!red_out, "head_bwd_apply_bnred: the fused reduction exists for K <= %d only", KMAX);
    return ubg_head_bwd_apply(dy, a, x, w, mean, rstd, gamma, dbeta, dgamma, dx, partial, P, K, dtype, stream);
  }
  const int grid = grid_for(P, 64 * 4, UB_STATS_ROWS);
  UB_CUDA(cudaMemsetAsync(partial, 0, sizeof(float) * UB_STATS_ROWS * (K * 64 + K), stream));
  if (red_out) {
    UB_CHECK_ARG(red_a && red_mean && red_rstd && dx, "head_bwd_apply_bnred: bad args");
    UB_CUDA(cudaMemsetAsync(red_out, 0, sizeof(float) * UB_STATS_ROWS * 128, stream));
    UB_DISPATCH_T(dtype, UB_DISPATCH_K(K, (head_bwd_apply_kernel<T, KK, true><<<grid, TPB, 0, stream>>>(
                                              dy, a, (const T*)x, w, mean, rstd, gamma, dbeta, dgamma, (T*)dx, partial, P, (const T*)red_a, red_mean,
                                              red_rstd, red_out))));
  } else {
    static int ppi = -1;
    if (ppi < 0) {
      const char* e = getenv("UB_HEAD_PPI");          // A/B switch: pixels per thread per iteration (2 = round-1 kernel)
      ppi = e ? atoi(e) : 4;
    }
    if (ppi == 4 && K <= 4)
      UB_DISPATCH_T(dtype, UB_DISPATCH_K(K, (head_bwd_apply_kernel<T, (KK <= 4 ? KK : 1), false, 4><<<grid, TPB, 0, stream>>>(
                                                dy, a, (const T*)x, w, mean, rstd, gamma, dbeta, dgamma, (T*)dx, partial, P, nullptr, nullptr, nullptr,
                                                nullptr))));
    else
      UB_DISPATCH_T(dtype, UB_DISPATCH_K(K, (head_bwd_apply_kernel<T, KK, false><<<grid, TPB, 0, stream>>>(
                                                dy, a, (const T*)x, w, mean, rstd, gamma, dbeta, dgamma, (T*)dx, partial, P, nullptr, nullptr, nullptr,
                                                nullptr))));
  }
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_head_bwd_apply(const float* dy, const float* a, const void* x, const float* w, const float* mean, const float* rstd,
                      const float* gamma, const float* dbeta, const float* dgamma, void* dx, float* partial, long long P, int K, int dtype,
                      cudaStream_t stream) {
  return head_bwd_apply_launch(dy, a, x, w, mean, rstd, gamma, dbeta, dgamma, dx, partial, P, K, dtype, nullptr, nullptr, nullptr, nullptr, stream);
}

int ub_head_bwd_apply_bnred(const float* dy, const float* a, const void* x, const float* w, const float* mean, const float* rstd,
                            const float* gamma, const float* dbeta, const float* dgamma, void* dx, float* partial, long long P, int K,
                            int dtype, const void* red_a, const float* red_mean, const float* red_rstd, float* red_partial,
                            cudaStream_t stream) {
  UB_CHECK_ARG(red_partial, "head_bwd_apply_bnred: red_partial is null");
  return head_bwd_apply_launch(dy, a, x, w, mean, rstd, gamma, dbeta, dgamma, dx, partial, P, K, dtype, red_a, red_mean, red_rstd, red_partial,
                               stream);
}

int ub_head_argmax(const void* x, const float* w, const float* b, const float* scale, const float* shift, int K, int ntiles, int h, int wd,
                   const int* geo, unsigned char* mask, long long mask_ld, float* softmax_out, int dtype, cudaStream_t stream) {
  UB_CHECK_ARG(x && w && b && scale && shift && (mask || softmax_out) && ntiles > 0 && h > 0 && wd > 0, "head_argmax: bad args");
  UB_CHECK_ARG(!mask || geo, "head_argmax: a mask needs the tile geometry");
  UB_CHECK_SHAPE(K >= 1 && K <= UB_MAX_CLASSES_ANY, "head_argmax: K");
  UB_CHECK_ARG(dtype == UB_BF16 || dtype == UB_F32, "head_argmax: bad dtype %d", dtype);
  if (K > KMAX) return ubg_head_argmax(x, w, b, scale, shift, K, ntiles, h, wd, geo, mask, mask_ld, softmax_out, dtype, stream);
  const long long P = (long long)h * wd;
  int gx = grid_for(P, 32 * 4, ub_num_sms() * 8 / (ntiles < 8 ? ntiles : 8));
  if (gx < 1) gx = 1;
  dim3 grid(gx, ntiles);
  UB_DISPATCH_T(dtype, UB_DISPATCH_K(K, (head_argmax_kernel<T, KK><<<grid, TPB, 0, stream>>>((const T*)x, w, b, scale, shift, h, wd,
                                                                                           reinterpret_cast<const TileGeo*>(geo), mask,
                                                                                           mask_ld, softmax_out))));
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_maxpool2x2_fwd(const void* y, void* pooled, int N, int H, int W, int C, int dtype, cudaStream_t stream) {
  UB_CHECK_ARG(y && pooled && N > 0, "maxpool2x2_fwd: bad args");
  UB_CHECK_SHAPE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool2x2_fwd: C %% 8, even H/W");
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
  UB_DISPATCH_T(dtype, (maxpool_fwd_kernel<T><<<grid_for(total, TPB, ub_num_sms() * 16), TPB, 0, stream>>>((const T*)y, (T*)pooled, N, H, W, C)));
  UB_LAUNCH_CHECK();
  return UB_OK;
}

}  // extern "C"
