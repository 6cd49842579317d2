// Implicit-GEMM on tcgen05/TMEM fed by TMA  (sm_100a).
//
//   D[pixel, col] = sum_{tap, src, c}  A_src[pixel + shift(tap), c] * Wt[col, (tap, src, c)]
//
// One kernel family serves four reference ops (all NHWC bf16, fp32 accumulate):
//   conv3x3 forward  (UNet/model.py:30-35)   9 taps, 1-2 concatenated sources (skip first, model.py:57), bias+ReLU,
//                                            per-channel sum / sum-of-squares partials for the following BatchNorm
//   conv3x3 dgrad    (autodiff of the above) same kernel with 180-degree-rotated, transposed weights; the output
//                                            channel range may be split over two tensors (gradient of a concat)
//   deconv2x2 forward (UNet/model.py:41-46)  1 tap, N = 4*Cout columns scattered through 4 strided output views
//   deconv2x2 dgrad                          4 strided source views, 1 tap
//
// M tile = 128 pixels = an 8x16 (h x w) patch of one image; the 'same' zero padding and ragged image edges come
// from TMA out-of-bounds zero fill / store clipping.  K is walked in 64-channel blocks (one 128-byte swizzle row).
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2-9 = epilogue (epilogue.cuh).
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
#include <stdlib.h>

#include "common.cuh"
#include "epilogue.cuh"

namespace {

constexpr int TILE_W = 16;
constexpr int TILE_H = 8;
constexpr int BLOCK_M = 128;
constexpr int A_BYTES = BLOCK_M * 128;  // 128 pixels x 64 bf16

struct IgemmFwdParams {
  CUtensorMap a_map[4];
  CUtensorMap b_map;
  CUtensorMap o_map[4];
  CUtensorMap red_map;  // RED: saved activation `a` of the BatchNorm'd tensor whose gradient a dgrad writes (epilogue.cuh)
  int nsrc;
  int cblk[4];       // 64-channel blocks per source
  int ntaps;
  int tap_dh[9];
  int tap_dw[9];
  int kb_per_tap;    // sum(cblk)
  int num_kb;        // ntaps * kb_per_tap
  int H, W;          // pixel space of the A sources
  int tiles_w, tiles_h;
  int n_tiles, total_tiles;
  int blocks_per_omap;
  EpiParams ep;       // bias (index = column % bias_mod), relu, stats [UB_STATS_ROWS][2][ncols]
  int ncols;
  const BnFin* fin;   // host-side only: BatchNorm to finalise in the last CTA (forward with statistics), or null
};

template <int BLOCK_N, int STAGES, int RED = 0>
struct SmemLayout {
  using E = EpiSmem<BLOCK_N, 1, RED>;
  static constexpr int B_BYTES = BLOCK_N * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OFF_EPI = STAGES * STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_EPI + E::TOTAL;             // full[S], empty[S], tfull[2], tempty[2]
  static constexpr int OFF_TMEM = OFF_BAR + (2 * STAGES + 4) * 8;
  static constexpr int TOTAL = OFF_TMEM + 16 + 1024;            // + alignment slack
};

template <int BLOCK_N, int STAGES, int RED = 0>
__global__ void __launch_bounds__(64 + EPI_THREADS, 1) igemm_fwd_kernel(const __grid_constant__ IgemmFwdParams p) {
  using L = SmemLayout<BLOCK_N, STAGES, RED>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::OFF_TMEM);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.nsrc; ++i) tma_prefetch_desc(&p.a_map[i]);
    tma_prefetch_desc(&p.b_map);
    tma_prefetch_desc(&p.o_map[0]);
    if (RED) tma_prefetch_desc(&p.red_map);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * BLOCK_N);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (warp == 0) {
    // ================= TMA producer (all lanes run the loop, one elected lane issues) =================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      const int img = m_tile / tiles_per_img;
      const int rem = m_tile - img * tiles_per_img;
      const int h0 = (rem / p.tiles_w) * TILE_H;
      const int w0 = (rem % p.tiles_w) * TILE_W;
      int kb = 0;
      for (int tap = 0; tap < p.ntaps; ++tap) {
        const int dh = p.tap_dh[tap], dw = p.tap_dw[tap];
        for (int src = 0; src < p.nsrc; ++src) {
          for (int cb = 0; cb < p.cblk[src]; ++cb, ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + stage * L::STAGE_BYTES;
            mbar_expect_tx_e(&full[stage], L::STAGE_BYTES);
            tma_load_4d_e(sa, &p.a_map[src], &full[stage], cb * 64, w0 + dw, h0 + dh, img);
            tma_load_2d_e(sa + A_BYTES, &p.b_map, &full[stage], kb * 64, n_tile * BLOCK_N);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer (all lanes run the loop, one elected lane issues) =================
    constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, 0, 0);
    const uint32_t tmem_u = warp_uniform(tmem_base);
    const uint32_t smem_base_u = warp_uniform(smem_u32(smem));
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_u + as * BLOCK_N;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_base_u + stage * L::STAGE_BYTES;
        const uint64_t adesc = make_smem_desc(sa, 16, 1024);
        const uint64_t bdesc = make_smem_desc(sa + A_BYTES, 16, 1024);
        tc_mma4_bf16_e(d_tmem, adesc, bdesc, idesc, kb != 0);   // 4 x (K = 16 bf16 = 32 bytes) per 64-channel block, one elect
        tc_commit_e(&empty[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      tc_commit_e(&tfull[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    __syncwarp();
  } else {
    // ================= epilogue (8 warps, epilogue.cuh) =================
    // host guarantees gridDim.x % n_tiles == 0, so this CTA's n_tile never changes
    const int n_tile = blockIdx.x % p.n_tiles;
    Epilogue<BLOCK_N, 1, TILE_W, RED> epi(smem + L::OFF_EPI, p.ep, tmem_base, tfull, tempty, threadIdx.x - 64, warp);
    epi.red_map = &p.red_map;
    epi.load_vectors(n_tile);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_tiles;
      const int img = m_tile / tiles_per_img;
      const int rem = m_tile - img * tiles_per_img;
      const int h0 = (rem / p.tiles_w) * TILE_H;
      const int w0 = (rem % p.tiles_w) * TILE_W;
      epi.tile(h0, w0, [&](const uint8_t* blk, int b) {
        const int j = n_tile * (BLOCK_N / 64) + b;
        const int map = j / p.blocks_per_omap;
        tma_store_4d(&p.o_map[map], blk, (j - map * p.blocks_per_omap) * 64, w0, h0, img);
      }, 0, true, true, BLOCK_N, img);
    }
    epi.finish(n_tile, blockIdx.x / p.n_tiles);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BLOCK_N);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct IgemmFwdLaunch {
  IgemmFwdParams p;
  int n_img;
  int ncols;
};

template <int BLOCK_N, int STAGES, int RED = 0>
int launch_t(IgemmFwdParams& p, int n_img, cudaStream_t stream) {
  using L = SmemLayout<BLOCK_N, STAGES, RED>;
  static_assert(L::TOTAL <= 232448, "smem budget");
  auto kern = igemm_fwd_kernel<BLOCK_N, STAGES, RED>;
  static bool attr_done = false;
  if (!attr_done) {
    UB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_done = true;
  }
  p.n_tiles = p.ncols / BLOCK_N;
  p.tiles_w = (p.W + TILE_W - 1) / TILE_W;
  p.tiles_h = (p.H + TILE_H - 1) / TILE_H;
  const long long m_tiles = (long long)n_img * p.tiles_w * p.tiles_h;
  const long long total = m_tiles * p.n_tiles;
  UB_CHECK_SHAPE(total > 0 && total < (1ll << 31), "igemm: tile count out of range");
  p.total_tiles = (int)total;
  int sms = ub_num_sms();
  int grid = (sms / p.n_tiles) * p.n_tiles;
  if (grid <= 0) grid = p.n_tiles;
  if (grid > p.total_tiles) grid = p.total_tiles;
  UB_CHECK_SHAPE(grid / p.n_tiles <= UB_STATS_ROWS, "igemm: stats rows");
  const BnFin* fin = p.fin;
  p.fin = nullptr;
  const bool fused = fin && p.ep.stats && epi_fin_bytes(p.ncols) <= L::E::OUT_BYTES;
  if (fused) epi_set_fin(p.ep, *fin, grid / p.n_tiles);
  else if (p.ep.stats) UB_CUDA(cudaMemsetAsync(p.ep.stats, 0, sizeof(float) * UB_STATS_ROWS * 2 * p.ncols, stream));
  if (RED) UB_CUDA(cudaMemsetAsync(p.ep.red_out, 0, sizeof(float) * UB_STATS_ROWS * 2 * p.ep.red_ncols, stream));
  kern<<<grid, 64 + EPI_THREADS, L::TOTAL, stream>>>(p);
  UB_LAUNCH_CHECK();
  if (fin && !fused)
    return ub_bn_finalize(p.ep.stats, p.ncols, fin->groups, (long long)fin->count, fin->mean, fin->rstd, fin->moving_mean, fin->moving_var,
                          fin->momentum, fin->eps, stream);
  return UB_OK;
}

int launch(IgemmFwdParams& p, int n_img, cudaStream_t stream) {
  p.kb_per_tap = 0;
  for (int i = 0; i < p.nsrc; ++i) p.kb_per_tap += p.cblk[i];
  p.num_kb = p.kb_per_tap * p.ntaps;
  p.ep.ncols = p.ncols;
  p.ep.H = p.H;
  p.ep.W = p.W;
  UB_CHECK_SHAPE(p.ncols % 64 == 0 && p.num_kb > 0, "igemm: columns must be a multiple of 64");
  if (p.ep.red_out) {          // a dgrad that also accumulates the BatchNorm-backward sums of the tensor it writes: room for the `a` tile
    if (p.ncols % 128 == 0) return launch_t<128, 4, 1>(p, n_img, stream);
    return launch_t<64, 8, 1>(p, n_img, stream);
  }
  if (p.ncols % 256 == 0) return launch_t<256, 3>(p, n_img, stream);
  if (p.ncols % 128 == 0) return launch_t<128, 5>(p, n_img, stream);
  return launch_t<64, 8>(p, n_img, stream);
}

void set_taps3x3(IgemmFwdParams& p) {
  p.ntaps = 9;
  for (int t = 0; t < 9; ++t) {
    p.tap_dh[t] = t / 3 - 1;
    p.tap_dw[t] = t % 3 - 1;
  }
}

int dense_map(CUtensorMap* m, const void* base, int C, int W, int H, int N) {
  return ub_tmap_act4d(m, base, C, W, H, N, (long long)C * 2, (long long)W * C * 2, (long long)H * W * C * 2, TILE_W, TILE_H);
}
// view of a [N, 2h, 2w, C] tensor restricted to rows 2i+a, cols 2j+b: dims {C, w, h, N}
int strided_map(CUtensorMap* m, const void* base, int C, int w, int h, int N, int a, int b) {
  const long long Wout = 2ll * w, Hout = 2ll * h;
  const uint8_t* pbase = reinterpret_cast<const uint8_t*>(base) + ((long long)a * Wout + b) * C * 2;
  return ub_tmap_act4d(m, pbase, C, w, h, N, 2ll * C * 2, 2ll * Wout * C * 2, Hout * Wout * C * 2, TILE_W, TILE_H);
}

}  // namespace

// halo-patch kernels (igemm_conv3.cu)
int ub_conv3_halo_fwd(const void* x0, int C0, const void* x1, int C1, const void* w, const float* bias, const float* post_scale,
                      const float* post_shift, void* out, float* stats, int N, int H, int W, int Cout, int relu, cudaStream_t stream,
                      int bias_cases = 0, const BnFin* fin = nullptr);
int ub_conv3_halo_dgrad(const void* dz, int Cout, const void* w_t, void* dx0, int C0, void* dx1, int C1, int N, int H, int W,
                        const void* red_a, const float* red_mean, const float* red_rstd, float* red_partial, cudaStream_t stream);
static bool legacy_conv3() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("UB_CONV3_LEGACY");      // A/B switch for tools/bench_layers.py: one TMA box per filter tap
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

extern "C" {

int ub_conv3x3_fwd(const void* x0, int C0, const void* x1, int C1, const void* w, const float* bias, void* out,
                   float* stats, int N, int H, int W, int Cout, int relu, cudaStream_t stream) {
  UB_CHECK_ARG(x0 && w && out, "conv3x3_fwd: null pointer");
  UB_CHECK_SHAPE(C0 > 0 && C0 % 64 == 0 && C1 >= 0 && C1 % 64 == 0 && Cout % 64 == 0 && (C1 == 0 || x1),
                 "conv3x3_fwd: channels must be multiples of 64 (C0=%d C1=%d Cout=%d)", C0, C1, Cout);
  UB_CHECK_SHAPE(N > 0 && H > 0 && W > 0, "conv3x3_fwd: bad N/H/W");
  if (!legacy_conv3()) return ub_conv3_halo_fwd(x0, C0, x1, C1, w, bias, nullptr, nullptr, out, stats, N, H, W, Cout, relu, stream);
  IgemmFwdParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = dense_map(&p.a_map[0], x0, C0, W, H, N))) return rc;
  p.nsrc = 1;
  p.cblk[0] = C0 / 64;
  if (C1 > 0) {
    if ((rc = dense_map(&p.a_map[1], x1, C1, W, H, N))) return rc;
    p.nsrc = 2;
    p.cblk[1] = C1 / 64;
  }
  const int Cin = C0 + C1;
  const int bn = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0 ? 128 : 64);
  if ((rc = ub_tmap_mat2d(&p.b_map, w, Cout, 9ll * Cin, bn))) return rc;
  if ((rc = dense_map(&p.o_map[0], out, Cout, W, H, N))) return rc;
  set_taps3x3(p);
  p.H = H;
  p.W = W;
  p.blocks_per_omap = Cout / 64;
  p.ep.bias = bias;
  p.ep.bias_mod = Cout;
  p.ep.relu = relu;
  p.ep.stats = stats;
  p.ncols = Cout;
  return launch(p, N, stream);
}

// Forward of a convolution whose INPUT is the pre-BatchNorm activation of its producer(s) (BatchNorm folded into the weights,
// ub_fold_conv3_weights): bias9 = [9][Cout], one bias vector per border case (row case * 3 + column case; 0 = first row /
// column, 1 = interior, 2 = last) because 'same' padding is applied AFTER BatchNorm (UNet/model.py:30-36).  H, W >= 2.
int ub_conv3x3_fwd_cases(const void* x0, int C0, const void* x1, int C1, const void* w, const float* bias9, void* out, float* stats, int N,
                         int H, int W, int Cout, int relu, cudaStream_t stream) {
  UB_CHECK_ARG(x0 && w && out && bias9, "conv3x3_fwd_cases: null pointer");
  UB_CHECK_SHAPE(C0 > 0 && C0 % 64 == 0 && C1 >= 0 && C1 % 64 == 0 && Cout % 64 == 0 && (C1 == 0 || x1),
                 "conv3x3_fwd_cases: channels must be multiples of 64 (C0=%d C1=%d Cout=%d)", C0, C1, Cout);
  UB_CHECK_SHAPE(N > 0 && H >= 2 && W >= 2, "conv3x3_fwd_cases: bad N/H/W");
  return ub_conv3_halo_fwd(x0, C0, x1, C1, w, bias9, nullptr, nullptr, out, stats, N, H, W, Cout, relu, stream, 1);
}

// Inference form: y = relu(conv + b) * scale + shift with the BatchNorm moving statistics folded into scale/shift
// (training=False, UNet/model.py:240 / inference.py:105): no separate normalisation pass.
int ub_conv3x3_fwd_affine(const void* x0, int C0, const void* x1, int C1, const void* w, const float* bias, const float* scale,
                          const float* shift, void* out, int N, int H, int W, int Cout, int relu, cudaStream_t stream) {
  UB_CHECK_ARG(x0 && w && out && scale && shift, "conv3x3_fwd_affine: null pointer");
  UB_CHECK_SHAPE(C0 > 0 && C0 % 64 == 0 && C1 >= 0 && C1 % 64 == 0 && Cout % 64 == 0 && (C1 == 0 || x1),
                 "conv3x3_fwd_affine: channels must be multiples of 64 (C0=%d C1=%d Cout=%d)", C0, C1, Cout);
  UB_CHECK_SHAPE(N > 0 && H > 0 && W > 0, "conv3x3_fwd_affine: bad N/H/W");
  return ub_conv3_halo_fwd(x0, C0, x1, C1, w, bias, scale, shift, out, nullptr, N, H, W, Cout, relu, stream);
}

int ub_conv3x3_dgrad(const void* dz, int Cout, const void* w_t, void* dx0, int C0, void* dx1, int C1, int N, int H,
                     int W, cudaStream_t stream) {
  UB_CHECK_ARG(dz && w_t && dx0, "conv3x3_dgrad: null pointer");
  UB_CHECK_SHAPE(Cout % 64 == 0 && C0 % 64 == 0 && C0 > 0 && (C1 == 0 || (C1 == C0 && dx1)),
                 "conv3x3_dgrad: channels must be multiples of 64 and split halves equal (Cout=%d C0=%d C1=%d)", Cout, C0, C1);
  if (!legacy_conv3()) return ub_conv3_halo_dgrad(dz, Cout, w_t, dx0, C0, dx1, C1, N, H, W, nullptr, nullptr, nullptr, nullptr, stream);
  IgemmFwdParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = dense_map(&p.a_map[0], dz, Cout, W, H, N))) return rc;
  p.nsrc = 1;
  p.cblk[0] = Cout / 64;
  const int Cin = C0 + C1;
  const int bn = (Cin % 256 == 0) ? 256 : (Cin % 128 == 0 ? 128 : 64);
  if ((rc = ub_tmap_mat2d(&p.b_map, w_t, Cin, 9ll * Cout, bn))) return rc;
  if ((rc = dense_map(&p.o_map[0], dx0, C0, W, H, N))) return rc;
  if (C1 > 0 && (rc = dense_map(&p.o_map[1], dx1, C1, W, H, N))) return rc;
  set_taps3x3(p);
  p.H = H;
  p.W = W;
  p.blocks_per_omap = C0 / 64;
  p.ep.bias = nullptr;
  p.ep.bias_mod = 1;
  p.ep.relu = 0;
  p.ep.stats = nullptr;
  p.ncols = Cin;
  return launch(p, N, stream);
}

static int deconv_fwd_impl(const void* x, int Cin, const void* w, const float* bias, const float* scale, const float* shift, void* out,
                           float* stats, int N, int h, int wd, int Cout, cudaStream_t stream, const BnFin* fin = nullptr);

/* dgrad fused with the backward-BatchNorm reduction of the tensor whose gradient it writes (the last output: dx1 of a
 * concat dgrad, else dx0): partial[UB_STATS_ROWS][2][C] = {sum dy, rstd * sum dy * (a - mean)} -- replaces ub_bn_bwd_reduce */
int ub_conv3x3_dgrad_bnred(const void* dz, int Cout, const void* w_t, void* dx0, int C0, void* dx1, int C1, int N, int H, int W,
                           const void* a, const float* mean, const float* rstd, float* partial, cudaStream_t stream) {
  UB_CHECK_ARG(dz && w_t && dx0 && a && mean && rstd && partial, "conv3x3_dgrad_bnred: null pointer");
  UB_CHECK_SHAPE(Cout % 64 == 0 && C0 % 64 == 0 && C0 > 0 && (C1 == 0 || (C1 == C0 && dx1)),
                 "conv3x3_dgrad_bnred: channels must be multiples of 64 and split halves equal (Cout=%d C0=%d C1=%d)", Cout, C0, C1);
  return ub_conv3_halo_dgrad(dz, Cout, w_t, dx0, C0, dx1, C1, N, H, W, a, mean, rstd, partial, stream);
}

int ub_deconv2x2_fwd(const void* x, int Cin, const void* w, const float* bias, void* out, float* stats, int N, int h,
                     int wd, int Cout, cudaStream_t stream) {
  return deconv_fwd_impl(x, Cin, w, bias, nullptr, nullptr, out, stats, N, h, wd, Cout, stream);
}

int ub_deconv2x2_fwd_affine(const void* x, int Cin, const void* w, const float* bias, const float* scale, const float* shift, void* out,
                            int N, int h, int wd, int Cout, cudaStream_t stream) {
  UB_CHECK_ARG(scale && shift, "deconv2x2_fwd_affine: null pointer");
  return deconv_fwd_impl(x, Cin, w, bias, scale, shift, out, nullptr, N, h, wd, Cout, stream);
}

static BnFin make_fin(float* mean, float* rstd, float* moving_mean, float* moving_var, double count, float momentum, float eps, int groups,
                      unsigned int* counter) {
  BnFin f;
  f.mean = mean;
  f.rstd = rstd;
  f.moving_mean = moving_mean;
  f.moving_var = moving_var;
  f.count = count;
  f.momentum = momentum;
  f.eps = eps;
  f.groups = groups;
  f.counter = counter;
  return f;
}

int ub_conv3x3_fwd_bn(const void* x0, int C0, const void* x1, int C1, const void* w, const float* bias, int bias_cases, void* out, float* stats,
                      int N, int H, int W, int Cout, int relu, float* mean, float* rstd, float* moving_mean, float* moving_var, float momentum,
                      float eps, unsigned int* counter, cudaStream_t stream) {
  UB_CHECK_ARG(x0 && w && out && stats && mean && rstd && counter && (!bias_cases || bias), "conv3x3_fwd_bn: null pointer");
  UB_CHECK_ARG((moving_mean == nullptr) == (moving_var == nullptr), "conv3x3_fwd_bn: moving_mean and moving_var go together");
  UB_CHECK_SHAPE(C0 > 0 && C0 % 64 == 0 && C1 >= 0 && C1 % 64 == 0 && Cout % 64 == 0 && (C1 == 0 || x1),
                 "conv3x3_fwd_bn: channels must be multiples of 64 (C0=%d C1=%d Cout=%d)", C0, C1, Cout);
  UB_CHECK_SHAPE(N > 0 && H > 0 && W > 0 && (!bias_cases || (H >= 2 && W >= 2)), "conv3x3_fwd_bn: bad N/H/W");
  const BnFin f = make_fin(mean, rstd, moving_mean, moving_var, (double)N * H * W, momentum, eps, 1, counter);
  return ub_conv3_halo_fwd(x0, C0, x1, C1, w, bias, nullptr, nullptr, out, stats, N, H, W, Cout, relu, stream, bias_cases ? 1 : 0, &f);
}

int ub_deconv2x2_fwd_bn(const void* x, int Cin, const void* w, const float* bias, void* out, float* stats, int N, int h, int wd, int Cout,
                        float* mean, float* rstd, float* moving_mean, float* moving_var, float momentum, float eps, unsigned int* counter,
                        cudaStream_t stream) {
  UB_CHECK_ARG(stats && mean && rstd && counter, "deconv2x2_fwd_bn: null pointer");
  UB_CHECK_ARG((moving_mean == nullptr) == (moving_var == nullptr), "deconv2x2_fwd_bn: moving_mean and moving_var go together");
  const BnFin f = make_fin(mean, rstd, moving_mean, moving_var, 4.0 * N * h * wd, momentum, eps, 4, counter);
  return deconv_fwd_impl(x, Cin, w, bias, nullptr, nullptr, out, stats, N, h, wd, Cout, stream, &f);
}

static int deconv_fwd_impl(const void* x, int Cin, const void* w, const float* bias, const float* scale, const float* shift, void* out,
                           float* stats, int N, int h, int wd, int Cout, cudaStream_t stream, const BnFin* fin) {
  UB_CHECK_ARG(x && w && out, "deconv2x2_fwd: null pointer");
  UB_CHECK_SHAPE(Cin % 64 == 0 && Cout % 64 == 0 && Cin > 0 && Cout > 0, "deconv2x2_fwd: channels must be multiples of 64");
  IgemmFwdParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = dense_map(&p.a_map[0], x, Cin, wd, h, N))) return rc;
  p.nsrc = 1;
  p.cblk[0] = Cin / 64;
  const int ncols = 4 * Cout;
  if ((rc = ub_tmap_mat2d(&p.b_map, w, ncols, Cin, ncols % 256 == 0 ? 256 : (ncols % 128 == 0 ? 128 : 64)))) return rc;
  for (int ab = 0; ab < 4; ++ab)
    if ((rc = strided_map(&p.o_map[ab], out, Cout, wd, h, N, ab >> 1, ab & 1))) return rc;
  p.ntaps = 1;
  p.H = h;
  p.W = wd;
  p.blocks_per_omap = Cout / 64;
  p.ep.bias = bias;
  p.ep.bias_mod = Cout;
  p.ep.post_scale = scale;
  p.ep.post_shift = shift;
  p.ep.relu = 0;
  p.ep.stats = stats;
  p.ncols = ncols;
  p.fin = fin;
  return launch(p, N, stream);
}

static int deconv_dgrad_impl(const void* dz, int Cout, const void* w_t, void* dx, int Cin, int N, int h, int wd, const void* red_a,
                             const float* red_mean, const float* red_rstd, float* red_partial, cudaStream_t stream) {
  UB_CHECK_ARG(dz && w_t && dx, "deconv2x2_dgrad: null pointer");
  UB_CHECK_SHAPE(Cin % 64 == 0 && Cout % 64 == 0 && Cin > 0 && Cout > 0, "deconv2x2_dgrad: channels must be multiples of 64");
  IgemmFwdParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  if (red_partial) {
    if ((rc = dense_map(&p.red_map, red_a, Cin, wd, h, N))) return rc;
    p.ep.red_mean = red_mean;
    p.ep.red_rstd = red_rstd;
    p.ep.red_out = red_partial;
    p.ep.red_blk_begin = 0;
    p.ep.red_ncols = Cin;
  }
  for (int ab = 0; ab < 4; ++ab) {
    if ((rc = strided_map(&p.a_map[ab], dz, Cout, wd, h, N, ab >> 1, ab & 1))) return rc;
    p.cblk[ab] = Cout / 64;
  }
  p.nsrc = 4;
  // the weight box must match the column tile launch() picks: 256 / 128 / 64, and never 256 with the fused reduction
  const int bn = (Cin % 256 == 0 && !red_partial) ? 256 : (Cin % 128 == 0 ? 128 : 64);
  if ((rc = ub_tmap_mat2d(&p.b_map, w_t, Cin, 4ll * Cout, bn))) return rc;
  if ((rc = dense_map(&p.o_map[0], dx, Cin, wd, h, N))) return rc;
  p.ntaps = 1;
  p.H = h;
  p.W = wd;
  p.blocks_per_omap = Cin / 64;
  p.ep.bias = nullptr;
  p.ep.bias_mod = 1;
  p.ep.relu = 0;
  p.ep.stats = nullptr;
  p.ncols = Cin;
  return launch(p, N, stream);
}

int ub_deconv2x2_dgrad(const void* dz, int Cout, const void* w_t, void* dx, int Cin, int N, int h, int wd, cudaStream_t stream) {
  return deconv_dgrad_impl(dz, Cout, w_t, dx, Cin, N, h, wd, nullptr, nullptr, nullptr, nullptr, stream);
}

/* transposed-convolution dgrad fused with the backward-BatchNorm reduction of the tensor whose gradient it writes (the conv block in front
 * of the up-sampling, UNet/model.py:97-131): partial[UB_STATS_ROWS][2][Cin] = {sum dy, rstd * sum dy * (a - mean)} -- replaces
 * ub_bn_bwd_reduce for that layer */
int ub_deconv2x2_dgrad_bnred(const void* dz, int Cout, const void* w_t, void* dx, int Cin, int N, int h, int wd, const void* a,
                             const float* mean, const float* rstd, float* partial, cudaStream_t stream) {
  UB_CHECK_ARG(a && mean && rstd && partial, "deconv2x2_dgrad_bnred: null pointer");
  return deconv_dgrad_impl(dz, Cout, w_t, dx, Cin, N, h, wd, a, mean, rstd, partial, stream);
}

}  // extern "C"
