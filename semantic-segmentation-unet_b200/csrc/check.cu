// fp32 "check mode" convolutions (north_star: "<=1e-4 in an fp32-accumulate check mode").
// Plain CUDA-core kernels over fp32 NHWC activations with the SAME weight packings and entry-point shapes as the
// tcgen05 path, so the whole graph can run in fp32 storage + fp32 FMA.  They are deliberately simple (one output
// element per thread): they exist to separate "bf16 rounding" from "wrong kernel" when a parity test fails, and to
// let tests reach the 1e-4 bar.  They are never selected for bf16 storage.
#include "common.cuh"

namespace {

// out[p][co] = act(b[co] + sum_t sum_ci x[p+t][ci] * w[co][t][ci]); x = concat(x0, x1) along channels
__global__ void check_conv3x3_fwd_kernel(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1, const float* __restrict__ w,
                                         const float* __restrict__ bias, float* __restrict__ out0, int Co0, float* __restrict__ out1, int N, int H,
                                         int W, int Cout, int relu) {
  const int Cin = C0 + C1;
  const long long total = (long long)N * H * W * Cout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const long long px = i / Cout;
    const int wv = (int)(px % W);
    const int h = (int)((px / W) % H);
    const int n = (int)(px / ((long long)W * H));
    float acc = bias ? bias[co] : 0.f;
    for (int t = 0; t < 9; ++t) {
      const int hh = h + t / 3 - 1, ww = wv + t % 3 - 1;
      if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
      const long long q = ((long long)n * H + hh) * W + ww;
      const float* wr = w + ((size_t)co * 9 + t) * Cin;
      const float* xa = x0 + q * C0;
      for (int c = 0; c < C0; ++c) acc += xa[c] * wr[c];
      if (C1 > 0) {
        const float* xb = x1 + q * C1;
        for (int c = 0; c < C1; ++c) acc += xb[c] * wr[C0 + c];
      }
    }
    if (relu) acc = fmaxf(acc, 0.f);
    if (co < Co0) out0[px * Co0 + co] = acc;
    else out1[px * (Cout - Co0) + (co - Co0)] = acc;
  }
}

// dw[co][t][ci] = sum_p dz[p][co] * x[p+t][ci]
__global__ void check_conv3x3_wgrad_kernel(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1, const float* __restrict__ dz,
                                           float* __restrict__ dw, int N, int H, int W, int Cout) {
  const int Cin = C0 + C1;
  const long long total = (long long)Cout * 9 * Cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const int t = (int)((i / Cin) % 9);
    const int co = (int)(i / (9ll * Cin));
    const float* xs = ci < C0 ? x0 : x1;
    const int Cs = ci < C0 ? C0 : C1;
    const int cs = ci < C0 ? ci : ci - C0;
    float acc = 0.f;
    for (int n = 0; n < N; ++n)
      for (int h = 0; h < H; ++h) {
        const int hh = h + t / 3 - 1;
        if (hh < 0 || hh >= H) continue;
        for (int wv = 0; wv < W; ++wv) {
          const int ww = wv + t % 3 - 1;
          if (ww < 0 || ww >= W) continue;
          acc += dz[(((long long)n * H + h) * W + wv) * Cout + co] * xs[(((long long)n * H + hh) * W + ww) * Cs + cs];
        }
      }
    dw[i] = acc;
  }
}

// out[n,2i+a,2j+b,co] = bias[co] + sum_ci x[n,i,j,ci] * w[(ab*Cout+co)][ci]
__global__ void check_deconv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out,
                                        int N, int h, int wd, int Cin, int Cout) {
  const long long total = (long long)N * 2 * h * 2 * wd * Cout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const long long px = i / Cout;
    const int ox = (int)(px % (2 * wd));
    const int oy = (int)((px / (2 * wd)) % (2 * h));
    const int n = (int)(px / (4ll * wd * h));
    const int ab = (oy & 1) * 2 + (ox & 1);
    const float* xr = x + (((long long)n * h + oy / 2) * wd + ox / 2) * Cin;
    const float* wr = w + ((size_t)ab * Cout + co) * Cin;
    float acc = bias ? bias[co] : 0.f;
    for (int c = 0; c < Cin; ++c) acc += xr[c] * wr[c];
    out[i] = acc;
  }
}
// dx[n,i,j,ci] = sum_ab sum_co dz[n,2i+a,2j+b,co] * w[(ab*Cout+co)][ci]
__global__ void check_deconv_dgrad_kernel(const float* __restrict__ dz, const float* __restrict__ w, float* __restrict__ dx, int N, int h, int wd,
                                          int Cin, int Cout) {
  const long long total = (long long)N * h * wd * Cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const long long px = i / Cin;
    const int x0 = (int)(px % wd);
    const int y0 = (int)((px / wd) % h);
    const int n = (int)(px / ((long long)wd * h));
    float acc = 0.f;
    for (int ab = 0; ab < 4; ++ab) {
      const float* dr = dz + (((long long)n * 2 * h + 2 * y0 + (ab >> 1)) * 2 * wd + 2 * x0 + (ab & 1)) * Cout;
      for (int co = 0; co < Cout; ++co) acc += dr[co] * w[((size_t)ab * Cout + co) * Cin + ci];
    }
    dx[i] = acc;
  }
}
// dw[(ab*Cout+co)][ci] = sum_p dz[n,2i+a,2j+b,co] * x[n,i,j,ci]
__global__ void check_deconv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dz, float* __restrict__ dw, int N, int h, int wd,
                                          int Cin, int Cout) {
  const long long total = 4ll * Cout * Cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const int co = (int)((i / Cin) % Cout);
    const int ab = (int)(i / ((long long)Cin * Cout));
    float acc = 0.f;
    for (int n = 0; n < N; ++n)
      for (int y0 = 0; y0 < h; ++y0)
        for (int x0 = 0; x0 < wd; ++x0)
          acc += dz[(((long long)n * 2 * h + 2 * y0 + (ab >> 1)) * 2 * wd + 2 * x0 + (ab & 1)) * Cout + co] *
                 x[(((long long)n * h + y0) * wd + x0) * Cin + ci];
    dw[i] = acc;
  }
}

inline int grid_of(long long total) {
  long long g = (total + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

extern "C" {

// Generic fp32 3x3 'same' convolution; also the fp32 dgrad when called with the rotated/transposed weight pack.
int ub_check_conv3x3(const float* x0, int C0, const float* x1, int C1, const float* w, const float* bias, float* out0, int Co0, float* out1,
                     int Co1, int N, int H, int W, int relu, cudaStream_t stream) {
  UB_CHECK_ARG(x0 && w && out0 && (C1 == 0 || x1) && (Co1 == 0 || out1), "check_conv3x3: bad args");
  const int Cout = Co0 + Co1;
  check_conv3x3_fwd_kernel<<<grid_of((long long)N * H * W * Cout), 256, 0, stream>>>(x0, C0, x1, C1, w, bias, out0, Co0, out1, N, H, W, Cout, relu);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_check_conv3x3_wgrad(const float* x0, int C0, const float* x1, int C1, const float* dz, int Cout, float* dw, int N, int H, int W,
                           cudaStream_t stream) {
  UB_CHECK_ARG(x0 && dz && dw && (C1 == 0 || x1), "check_conv3x3_wgrad: bad args");
  check_conv3x3_wgrad_kernel<<<grid_of((long long)Cout * 9 * (C0 + C1)), 256, 0, stream>>>(x0, C0, x1, C1, dz, dw, N, H, W, Cout);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_check_deconv2x2_fwd(const float* x, const float* w, const float* bias, float* out, int N, int h, int wd, int Cin, int Cout,
                           cudaStream_t stream) {
  UB_CHECK_ARG(x && w && out, "check_deconv2x2_fwd: bad args");
  check_deconv_fwd_kernel<<<grid_of((long long)N * 4 * h * wd * Cout), 256, 0, stream>>>(x, w, bias, out, N, h, wd, Cin, Cout);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_check_deconv2x2_dgrad(const float* dz, const float* w, float* dx, int N, int h, int wd, int Cin, int Cout, cudaStream_t stream) {
  UB_CHECK_ARG(dz && w && dx, "check_deconv2x2_dgrad: bad args");
  check_deconv_dgrad_kernel<<<grid_of((long long)N * h * wd * Cin), 256, 0, stream>>>(dz, w, dx, N, h, wd, Cin, Cout);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int ub_check_deconv2x2_wgrad(const float* x, const float* dz, float* dw, int N, int h, int wd, int Cin, int Cout, cudaStream_t stream) {
  UB_CHECK_ARG(x && dz && dw, "check_deconv2x2_wgrad: bad args");
  check_deconv_wgrad_kernel<<<grid_of(4ll * Cout * Cin), 256, 0, stream>>>(x, dz, dw, N, h, wd, Cin, Cout);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

}  // extern "C"
