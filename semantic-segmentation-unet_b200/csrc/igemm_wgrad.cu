// Weight-gradient GEMM on tcgen05/TMEM fed by TMA (sm_100a):   D[m, n] = sum_pixels  A[pixel(+shift), m] * B[pixel, n]
//
//   conv3x3 wgrad  (autodiff of UNet/model.py:30-35):  m = (tap, concat-source, ci), A = layer input x shifted by tap,
//                                                      n = co,                      B = dz
//   deconv2x2 wgrad (autodiff of UNet/model.py:41-46): m = ci, A = x ;  n = (a,b,co), B = dz viewed at rows 2i+a, cols 2j+b
//
// Both operands are NHWC activations, i.e. the reduction dimension (pixels) is the strided one: they are fed to the
// tensor core as MN-major SWIZZLE_128B tiles -- the smem image a {64ch x 16w x 4h} TMA box produces is exactly the
// canonical MN-major layout (128-byte rows = 64 channels of one pixel, 8-pixel swizzle atoms).
// Each CTA owns one 128 x BLOCK_N output tile and a contiguous range of 64-pixel patches (split-K); partial tiles go to
// a fp32 workspace [split][m][n] and ub_wgrad_reduce sums them in a fixed order (deterministic) into dW[n][m].
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int PATCH_W = 16;
constexpr int PATCH_H = 4;
constexpr int BOX_BYTES = 64 * 128;  // 64 pixels x 64 channels bf16

struct SideDesc {
  CUtensorMap map[4];
  int nmaps;
  int cblk[4];     // 64-channel blocks per map
  int ntaps;
  int tap_dh[9];
  int tap_dw[9];
  int per_tap;     // sum(cblk)
  int nblocks;     // ntaps * per_tap
};

struct WgradParams {
  SideDesc a, b;
  int H, W;                 // pixel space
  int patches_w, patches_h, total_patches;
  int m_tiles, n_tiles, splits;
  int mrows, ncols;         // valid rows (a.nblocks*64) / columns (b.nblocks*64)
  float* ws;                // [splits][mrows][ncols]
};

__device__ __forceinline__ void decode_block(const SideDesc& s, int blk, int& map, int& c0, int& dh, int& dw) {
  const int tap = blk / s.per_tap;
  int r = blk - tap * s.per_tap;
  map = 0;
  while (r >= s.cblk[map]) {
    r -= s.cblk[map];
    ++map;
  }
  c0 = r * 64;
  dh = s.tap_dh[tap];
  dw = s.tap_dw[tap];
}

template <int BLOCK_N, int STAGES>
struct WgSmem {
  static constexpr int A_BYTES = 2 * BOX_BYTES;
  static constexpr int B_BYTES = (BLOCK_N / 64) * BOX_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OFF_BAR = STAGES * STAGE_BYTES;
  static constexpr int OFF_TMEM = OFF_BAR + (2 * STAGES + 1) * 8;
  static constexpr int TOTAL = OFF_TMEM + 16 + 1024;
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(192, 1) igemm_wgrad_kernel(const __grid_constant__ WgradParams p) {
  using L = WgSmem<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::OFF_TMEM);
  constexpr uint32_t TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work unit: (m_tile, n_tile, split); split fastest so neighbouring CTAs stream different pixels
  const int split = blockIdx.x % p.splits;
  const int tile = blockIdx.x / p.splits;
  const int n_tile = tile % p.n_tiles;
  const int m_tile = tile / p.n_tiles;
  const int p_begin = (int)((long long)p.total_patches * split / p.splits);
  const int p_end = (int)((long long)p.total_patches * (split + 1) / p.splits);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tfull, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // TMA producer: all lanes run the loop, one elected lane issues (see common.cuh)
    // M side: two 64-row blocks (the second clamps to the last valid block when the block count is odd)
    int amap[2], ac0[2], adh[2], adw[2];
    for (int i = 0; i < 2; ++i) {
      int blk = m_tile * 2 + i;
      if (blk >= p.a.nblocks) blk = p.a.nblocks - 1;
      decode_block(p.a, blk, amap[i], ac0[i], adh[i], adw[i]);
    }
    int bmap[BLOCK_N / 64], bc0[BLOCK_N / 64], bdh[BLOCK_N / 64], bdw[BLOCK_N / 64];
#pragma unroll
    for (int j = 0; j < BLOCK_N / 64; ++j) decode_block(p.b, n_tile * (BLOCK_N / 64) + j, bmap[j], bc0[j], bdh[j], bdw[j]);
    const int per_img = p.patches_w * p.patches_h;
    int stage = 0;
    uint32_t phase = 0;
    for (int pt = p_begin; pt < p_end; ++pt) {
      const int img = pt / per_img;
      const int rem = pt - img * per_img;
      const int h0 = (rem / p.patches_w) * PATCH_H;
      const int w0 = (rem % p.patches_w) * PATCH_W;
      mbar_wait(&empty[stage], phase ^ 1);
      uint8_t* sa = smem + stage * L::STAGE_BYTES;
      mbar_expect_tx_e(&full[stage], L::STAGE_BYTES);
#pragma unroll
      for (int i = 0; i < 2; ++i)
        tma_load_4d_e(sa + i * BOX_BYTES, &p.a.map[amap[i]], &full[stage], ac0[i], w0 + adw[i], h0 + adh[i], img);
#pragma unroll
      for (int j = 0; j < BLOCK_N / 64; ++j)
        tma_load_4d_e(sa + L::A_BYTES + j * BOX_BYTES, &p.b.map[bmap[j]], &full[stage], bc0[j], w0 + bdw[j], h0 + bdh[j], img);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 1, 1);  // both operands MN-major
    const uint32_t tmem_u = warp_uniform(tmem_base);
    const uint32_t smem_base_u = warp_uniform(smem_u32(smem));
    int stage = 0;
    uint32_t phase = 0;
    for (int pt = p_begin; pt < p_end; ++pt) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint32_t sa = smem_base_u + stage * L::STAGE_BYTES;
      // 4 x 16 pixels, one elected asm block
      tc_mma_steps_bf16_e<4>(tmem_u, make_smem_desc(sa, BOX_BYTES, 1024), 2048 >> 4, make_smem_desc(sa + L::A_BYTES, BOX_BYTES, 1024), 2048 >> 4,
                             idesc, pt > p_begin);
      tc_commit_e(&empty[stage]);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    tc_commit_e(tfull);
    __syncwarp();
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int m = m_tile * 128 + row;
    float* dst = p.ws + ((size_t)split * p.mrows + m) * p.ncols + n_tile * BLOCK_N;
    if (p_end > p_begin) {
      mbar_wait(tfull, 0);
      tc_fence_after();
    }
#pragma unroll
    for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
      uint32_t v[32];
      if (p_end > p_begin) {
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + chunk * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (m < p.mrows) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(dst + chunk * 32 + q * 4) = make_uint4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// dW[n][m] = sum_s ws[s][m][n]   (fixed summation order; 32x32 transpose through smem)
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int splits, int mrows, int ncols) {
  __shared__ float t[32][33];
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int m = m0 + r, n = n0 + tx;
    float acc = 0.f;
    if (m < mrows && n < ncols)
      for (int s = 0; s < splits; ++s) acc += ws[((size_t)s * mrows + m) * ncols + n];
    t[r][tx] = acc;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, m = m0 + tx;
    if (m < mrows && n < ncols) dw[(size_t)n * mrows + m] = t[tx][r];
  }
}

template <int BLOCK_N, int STAGES>
int launch_wg(WgradParams& p, cudaStream_t stream) {
  using L = WgSmem<BLOCK_N, STAGES>;
  static_assert(L::TOTAL <= 232448, "smem budget");
  auto kern = igemm_wgrad_kernel<BLOCK_N, STAGES>;
  static bool attr_done = false;
  if (!attr_done) {
    UB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_done = true;
  }
  kern<<<p.m_tiles * p.n_tiles * p.splits, 192, L::TOTAL, stream>>>(p);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int pick_block_n(int ncols) { return ncols % 256 == 0 ? 256 : (ncols % 128 == 0 ? 128 : 64); }

void finish_side(SideDesc& s) {
  s.per_tap = 0;
  for (int i = 0; i < s.nmaps; ++i) s.per_tap += s.cblk[i];
  s.nblocks = s.per_tap * s.ntaps;
}

int plan(WgradParams& p, int n_img, int H, int W) {
  finish_side(p.a);
  finish_side(p.b);
  p.H = H;
  p.W = W;
  p.patches_w = (W + PATCH_W - 1) / PATCH_W;
  p.patches_h = (H + PATCH_H - 1) / PATCH_H;
  const long long tp = (long long)n_img * p.patches_w * p.patches_h;
  UB_CHECK_SHAPE(tp > 0 && tp < (1ll << 31), "wgrad: patch count out of range");
  p.total_patches = (int)tp;
  p.mrows = p.a.nblocks * 64;
  p.ncols = p.b.nblocks * 64;
  const int bn = pick_block_n(p.ncols);
  p.m_tiles = (p.a.nblocks + 1) / 2;
  p.n_tiles = p.ncols / bn;
  const int tiles = p.m_tiles * p.n_tiles;
  // split-K so the grid covers ~2 waves of SMs, at least 8 patches (512 pixels) per CTA
  int splits = (2 * ub_num_sms() + tiles - 1) / tiles;
  const int max_splits = (p.total_patches + 7) / 8;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.splits = splits;
  return UB_OK;
}

int run(WgradParams& p, float* dw, float* ws, size_t ws_bytes, cudaStream_t stream) {
  const size_t need = (size_t)p.splits * p.mrows * p.ncols * sizeof(float);
  UB_CHECK_ARG(ws && ws_bytes >= need, "wgrad: workspace too small (%zu < %zu)", ws_bytes, need);
  p.ws = ws;
  int rc;
  const int bn = pick_block_n(p.ncols);
  if (bn == 256) rc = launch_wg<256, 4>(p, stream);
  else if (bn == 128) rc = launch_wg<128, 6>(p, stream);
  else rc = launch_wg<64, 8>(p, stream);
  if (rc) return rc;
  dim3 grid((p.ncols + 31) / 32, (p.mrows + 31) / 32), block(32, 8);
  wgrad_reduce_kernel<<<grid, block, 0, stream>>>(ws, dw, p.splits, p.mrows, p.ncols);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int map_dense(CUtensorMap* m, const void* base, int C, int W, int H, int N) {
  return ub_tmap_act4d(m, base, C, W, H, N, (long long)C * 2, (long long)W * C * 2, (long long)H * W * C * 2, PATCH_W, PATCH_H);
}
int map_strided(CUtensorMap* m, const void* base, int C, int w, int h, int N, int a, int b) {
  const long long Wout = 2ll * w, Hout = 2ll * h;
  const uint8_t* pbase = reinterpret_cast<const uint8_t*>(base) + ((long long)a * Wout + b) * C * 2;
  return ub_tmap_act4d(m, pbase, C, w, h, N, 2ll * C * 2, 2ll * Wout * C * 2, Hout * Wout * C * 2, PATCH_W, PATCH_H);
}

int build_conv(WgradParams& p, const void* x0, int C0, const void* x1, int C1, const void* dz, int Cout, int N, int H, int W) {
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = map_dense(&p.a.map[0], x0, C0, W, H, N))) return rc;
  p.a.nmaps = 1;
  p.a.cblk[0] = C0 / 64;
  if (C1 > 0) {
    if ((rc = map_dense(&p.a.map[1], x1, C1, W, H, N))) return rc;
    p.a.nmaps = 2;
    p.a.cblk[1] = C1 / 64;
  }
  p.a.ntaps = 9;
  for (int t = 0; t < 9; ++t) {
    p.a.tap_dh[t] = t / 3 - 1;
    p.a.tap_dw[t] = t % 3 - 1;
  }
  if ((rc = map_dense(&p.b.map[0], dz, Cout, W, H, N))) return rc;
  p.b.nmaps = 1;
  p.b.cblk[0] = Cout / 64;
  p.b.ntaps = 1;
  return plan(p, N, H, W);
}

int build_deconv(WgradParams& p, const void* x, int Cin, const void* dz, int Cout, int N, int h, int w) {
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = map_dense(&p.a.map[0], x, Cin, w, h, N))) return rc;
  p.a.nmaps = 1;
  p.a.cblk[0] = Cin / 64;
  p.a.ntaps = 1;
  for (int ab = 0; ab < 4; ++ab) {
    if ((rc = map_strided(&p.b.map[ab], dz, Cout, w, h, N, ab >> 1, ab & 1))) return rc;
    p.b.cblk[ab] = Cout / 64;
  }
  p.b.nmaps = 4;
  p.b.ntaps = 1;
  return plan(p, N, h, w);
}

}  // namespace

// halo-patch kernel (igemm_wgrad_halo.cu): the product path for conv3x3; this file's kernel serves the deconvs
long long ub_wgrad_halo_workspace_bytes(int C0, int C1, int Cout, int N, int H, int W);
int ub_wgrad_halo(const void* x0, int C0, const void* x1, int C1, const void* dz, int Cout, float* dw, void* workspace,
                  long long workspace_bytes, int N, int H, int W, cudaStream_t stream);
static bool legacy_wgrad() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("UB_WGRAD_LEGACY");      // A/B switch for tools/bench_layers.py: one TMA box per (tap, channel block)
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

extern "C" {

long long ub_conv3x3_wgrad_workspace_bytes(int C0, int C1, int Cout, int N, int H, int W) {
  if (!legacy_wgrad()) return ub_wgrad_halo_workspace_bytes(C0, C1, Cout, N, H, W);
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.a.nmaps = C1 > 0 ? 2 : 1;
  p.a.cblk[0] = C0 / 64;
  p.a.cblk[1] = C1 / 64;
  p.a.ntaps = 9;
  p.b.nmaps = 1;
  p.b.cblk[0] = Cout / 64;
  p.b.ntaps = 1;
  if (plan(p, N, H, W)) return -1;
  return (long long)p.splits * p.mrows * p.ncols * 4;
}

int ub_conv3x3_wgrad(const void* x0, int C0, const void* x1, int C1, const void* dz, int Cout, float* dw, void* workspace,
                     long long workspace_bytes, int N, int H, int W, cudaStream_t stream) {
  UB_CHECK_ARG(x0 && dz && dw, "conv3x3_wgrad: null pointer");
  UB_CHECK_SHAPE(C0 > 0 && C0 % 64 == 0 && C1 >= 0 && C1 % 64 == 0 && Cout % 64 == 0 && (C1 == 0 || x1),
                 "conv3x3_wgrad: channels must be multiples of 64 (C0=%d C1=%d Cout=%d)", C0, C1, Cout);
  UB_CHECK_SHAPE(N > 0 && H > 0 && W > 0, "conv3x3_wgrad: bad N/H/W");
  if (!legacy_wgrad()) return ub_wgrad_halo(x0, C0, x1, C1, dz, Cout, dw, workspace, workspace_bytes, N, H, W, stream);
  WgradParams p;
  int rc = build_conv(p, x0, C0, x1, C1, dz, Cout, N, H, W);
  if (rc) return rc;
  return run(p, dw, reinterpret_cast<float*>(workspace), (size_t)workspace_bytes, stream);
}

long long ub_deconv2x2_wgrad_workspace_bytes(int Cin, int Cout, int N, int h, int w) {
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.a.nmaps = 1;
  p.a.cblk[0] = Cin / 64;
  p.a.ntaps = 1;
  p.b.nmaps = 4;
  for (int i = 0; i < 4; ++i) p.b.cblk[i] = Cout / 64;
  p.b.ntaps = 1;
  if (plan(p, N, h, w)) return -1;
  return (long long)p.splits * p.mrows * p.ncols * 4;
}

int ub_deconv2x2_wgrad(const void* x, int Cin, const void* dz, int Cout, float* dw, void* workspace,
                       long long workspace_bytes, int N, int h, int w, cudaStream_t stream) {
  UB_CHECK_ARG(x && dz && dw, "deconv2x2_wgrad: null pointer");
  UB_CHECK_SHAPE(Cin % 128 == 0 && Cout % 64 == 0, "deconv2x2_wgrad: Cin must be a multiple of 128, Cout of 64");
  WgradParams p;
  int rc = build_deconv(p, x, Cin, dz, Cout, N, h, w);
  if (rc) return rc;
  return run(p, dw, reinterpret_cast<float*>(workspace), (size_t)workspace_bytes, stream);
}

}  // extern "C"
