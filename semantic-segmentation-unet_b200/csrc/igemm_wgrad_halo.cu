// conv3x3 weight gradient as a halo-patch implicit GEMM on tcgen05/TMEM (sm_100a).
//
//   dW[co][tap][ci] = sum_pixels  x[pixel + shift(tap), ci] * dz[pixel, co]               (autodiff of UNet/model.py:30-35)
//
// The reduction runs over pixels, so both operands are MN-major SWIZZLE_128B tiles (128-byte row = 64 channels of one
// pixel).  The first wgrad kernel (igemm_wgrad.cu) loaded one x box per (tap, channel block): (128 + N) * 128 bytes
// of shared-memory fill per 64 pixels, 96-128 B/clk/SM against the ~43 B/clk/SM the L2 can deliver -> 30-60 % of the
// tensor peak and worse on the 64/128-channel layers.  Here a CTA loads ONE halo'd x patch per 64-channel block and
// 16x8-pixel tile and reads the filter taps as shifted views of it (the XOR swizzle is a function of the absolute
// shared-memory address for MN-major operands too: tools/desc_probe.py), with one TMEM accumulator per tap:
//
//   mode P (>= 2 channel blocks, Cout % 128 == 0): M = 128 = two channel blocks (two patches, LBO = patch pitch),
//           the three taps of one filter row -> 3 accumulators x 128 columns; 72 KB per 1536 MMA clocks = 47 B/clk.
//   mode Q (any Cin/Cout multiple of 64):          M = 128 = two TAPS of one channel block (LBO = the address
//           distance between the two shifted views), all nine taps as five pairs -> 5 accumulators x 64 columns;
//           39 KB per 1920 MMA clocks = 20 B/clk (N = 64 MMAs are shared-memory-read bound at 2/3 of the peak).
//
// Work item = (channel block [pair], filter row | -, 64/128-column block of Cout); the pixel tiles of an item are split
// over `splits` CTAs; every CTA writes its accumulators in the final dW layout (a TMEM lane is a ci index, so a
// warp's store of one column is 128 contiguous bytes) either straight into dW (splits == 1) or into its slice of a
// [splits][Cout*9*Cin] workspace that ub_wgrad_sum_splits adds up in a fixed order (deterministic, no atomics).
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2-5 = epilogue.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int TW = 8, TH = 16;
constexpr int PW = TW + 2;
constexpr int DZ_BLK_BYTES = TW * TH * 128;            // 16384: 128 pixels x 64 channels

struct WgHaloParams {
  CUtensorMap x_map[2];        // box {64, PW, PH_MODE}
  CUtensorMap dz_map;          // box {64, TW, TH}
  int cblk0;                   // 64-channel blocks in source 0 (blocks >= cblk0 come from source 1)
  int cblk_total;
  int Cin, Cout;
  int tiles_w, tiles_h, total_tiles;
  int items, splits;
  int n_tiles;                 // Cout / BLOCK_N
  float* out;                  // dW (splits == 1) or workspace [splits][Cout*9*Cin]
  long long split_stride;      // elements between split slices
};

template <int MODE>
struct Cfg;
template <>
struct Cfg<0> {   // mode P
  static constexpr int BLOCK_N = 128;
  static constexpr int PH = TH;                          // filter row folded into the TMA coordinate
  static constexpr int PATCH_BYTES = PW * PH * 128;      // 20480
  static constexpr int PATCH_STRIDE = PATCH_BYTES;       // 1024-aligned already
  static constexpr int NPATCH = 2;
  static constexpr int NACC = 3;
  static constexpr int STAGES = 3;
};
template <>
struct Cfg<2> {   // mode P with a two-deep pipeline: 148 KB of shared memory, leaves room for co-resident memory-bound CTAs
  static constexpr int BLOCK_N = 128;
  static constexpr int PH = TH;
  static constexpr int PATCH_BYTES = PW * PH * 128;
  static constexpr int PATCH_STRIDE = PATCH_BYTES;
  static constexpr int NPATCH = 2;
  static constexpr int NACC = 3;
  static constexpr int STAGES = 2;
};
template <>
struct Cfg<1> {   // mode Q
  static constexpr int BLOCK_N = 64;
  static constexpr int PH = TH + 2;
  static constexpr int PATCH_BYTES = PW * PH * 128;      // 23040
  static constexpr int PATCH_STRIDE = 23 * 1024;
  static constexpr int NPATCH = 1;
  static constexpr int NACC = 5;
  static constexpr int STAGES = 5;
};

template <int MODE>
struct WhSmem {
  using C = Cfg<MODE>;
  static constexpr int DZ_BYTES = (C::BLOCK_N / 64) * DZ_BLK_BYTES;
  static constexpr int OFF_DZ = C::NPATCH * C::PATCH_STRIDE;
  static constexpr int STAGE_BYTES = OFF_DZ + DZ_BYTES;
  static constexpr int TX_BYTES = C::NPATCH * C::PATCH_BYTES + DZ_BYTES;
  static constexpr int OFF_BAR = C::STAGES * STAGE_BYTES;
  static constexpr int OFF_TMEM = OFF_BAR + (2 * C::STAGES + 1) * 8;
  static constexpr int TOTAL = OFF_TMEM + 16 + 1024;
  static_assert(STAGE_BYTES % 1024 == 0, "stage alignment");
};

// mode Q tap pairs: first/second tap of accumulator j (tap = 3 * dh + dw); accumulator 4's first half duplicates
// tap 5 and is not stored.
__device__ __constant__ int c_pair_a[5] = {0, 3, 6, 2, 5};
__device__ __constant__ int c_pair_b[5] = {1, 4, 7, 5, 8};

template <int MODE>
__global__ void __launch_bounds__(192, 1) wgrad_halo_kernel(const __grid_constant__ WgHaloParams p) {
  using C = Cfg<MODE>;
  using L = WhSmem<MODE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tfull = empty + C::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::OFF_TMEM);
  constexpr uint32_t TMEM_COLS = 512;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work unit: split fastest, so CTAs that run together stream different pixel ranges of the same operands
  const int split = blockIdx.x % p.splits;
  int item = blockIdx.x / p.splits;
  const int n_tile = item % p.n_tiles;
  item /= p.n_tiles;
  int dh = 0, cb0;
  if (MODE != 1) {
    dh = item % 3;
    cb0 = (item / 3) * 2;
  } else {
    cb0 = item;
  }
  const int t_begin = (int)((long long)p.total_tiles * split / p.splits);
  const int t_end = (int)((long long)p.total_tiles * (split + 1) / p.splits);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.x_map[0]);
    tma_prefetch_desc(&p.dz_map);
    for (int s = 0; s < C::STAGES; ++s) {
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
    // ================= TMA producer =================
    const int tiles_per_img = p.tiles_w * p.tiles_h;
    int stage = 0;
    uint32_t phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int img = t / tiles_per_img;
      const int rem = t - img * tiles_per_img;
      const int h0 = (rem / p.tiles_w) * TH;
      const int w0 = (rem % p.tiles_w) * TW;
      mbar_wait(&empty[stage], phase ^ 1);
      uint8_t* sa = smem + stage * L::STAGE_BYTES;
      mbar_expect_tx_e(&full[stage], L::TX_BYTES);
#pragma unroll
      for (int i = 0; i < C::NPATCH; ++i) {
        const int cb = cb0 + i;
        const int src = cb >= p.cblk0 ? 1 : 0;
        const int c0 = (src ? cb - p.cblk0 : cb) * 64;
        tma_load_4d_e(sa + i * C::PATCH_STRIDE, &p.x_map[src], &full[stage], c0, w0 - 1, h0 - 1 + dh, img);
      }
#pragma unroll
      for (int j = 0; j < C::BLOCK_N / 64; ++j)
        tma_load_4d_e(sa + L::OFF_DZ + j * DZ_BLK_BYTES, &p.dz_map, &full[stage], n_tile * C::BLOCK_N + j * 64, w0, h0, img);
      if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = make_idesc_bf16(128, C::BLOCK_N, 1, 1);     // both operands MN-major
    const uint32_t tmem_u = warp_uniform(tmem_base);
    const uint32_t smem_base_u = warp_uniform(smem_u32(smem));
    int stage = 0;
    uint32_t phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint32_t sa = smem_base_u + stage * L::STAGE_BYTES;
      const uint32_t acc_on = (t > t_begin) ? 1u : 0u;
      // K = 16 pixels = two tile rows per MMA; 8 K steps per accumulator in one elected asm block (the issue stream of this
      // warp, not the tensor pipe, was the limit with one elect per MMA)
      const uint64_t bdesc = make_smem_desc(sa + L::OFF_DZ, DZ_BLK_BYTES, TW * 128);
      constexpr uint64_t B_STEP = (2 * TW * 128) >> 4, A_STEP = (2 * PW * 128) >> 4;
      if (MODE != 1) {
        const uint64_t adesc0 = make_smem_desc(sa, C::PATCH_STRIDE, PW * 128);
#pragma unroll
        for (int dw = 0; dw < 3; ++dw)
          tc_mma_steps_bf16_e<8>(tmem_u + dw * C::BLOCK_N, adesc0 + (uint64_t)((dw * 128) >> 4), A_STEP, bdesc, B_STEP, idesc, acc_on);
      } else {
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          // pairs (0,1) (3,4) (6,7): second view one pixel to the right (LBO 128 B); (2,5) (5,8): one patch row down (LBO 1280 B)
          const int ta = (j < 3) ? 3 * j : (j == 3 ? 2 : 5);
          const int adh = ta / 3, adw = ta % 3;
          const uint32_t lbo = (j < 3) ? 128u : (uint32_t)(PW * 128);
          const uint64_t adesc = make_smem_desc(sa + (adh * PW + adw) * 128, lbo, PW * 128);
          tc_mma_steps_bf16_e<8>(tmem_u + j * C::BLOCK_N, adesc, A_STEP, bdesc, B_STEP, idesc, acc_on);
        }
      }
      tc_commit_e(&empty[stage]);
      if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
    }
    tc_commit_e(tfull);
    __syncwarp();
  } else {
    // ================= epilogue: TMEM -> dW layout =================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int half = row >> 6;
    const bool have = t_end > t_begin;
    if (have) {
      mbar_wait(tfull, 0);
      tc_fence_after();
    }
    float* out = p.out + (size_t)split * p.split_stride;
#pragma unroll 1
    for (int a = 0; a < C::NACC; ++a) {
      int tap, ci;
      bool store = true;
      if (MODE != 1) {
        tap = dh * 3 + a;
        ci = (cb0 + half) * 64 + (row & 63);
      } else {
        tap = half ? c_pair_b[a] : c_pair_a[a];
        ci = cb0 * 64 + (row & 63);
        store = !(a == 4 && half == 0);
      }
#pragma unroll 1
      for (int chunk = 0; chunk < C::BLOCK_N / 32; ++chunk) {
        uint32_t v[32];
        if (have) {
          tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + a * C::BLOCK_N + chunk * 32, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        if (store) {
          const int co0 = n_tile * C::BLOCK_N + chunk * 32;
          float* dst = out + ((size_t)co0 * 9 + tap) * p.Cin + ci;
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[(size_t)j * 9 * p.Cin] = __uint_as_float(v[j]);   // a warp writes 32 consecutive ci
        }
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

// dw[i] = sum_s ws[s][i]  (fixed order)
__global__ void __launch_bounds__(256) wgrad_sum_splits_kernel(const float4* __restrict__ ws, float4* __restrict__ dw, int splits, long long n4) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 acc = __ldg(ws + i);
    for (int s = 1; s < splits; ++s) {
      const float4 v = __ldg(ws + (size_t)s * n4 + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    dw[i] = acc;
  }
}

int pick_mode(int C0, int C1, int Cout) {
  const int cblk = (C0 + C1) / 64;
  return (cblk % 2 == 0 && (C0 / 64) % 2 == 0 && Cout % 128 == 0) ? 0 : 1;
}

// splits: fill whole waves of the 148 SMs; at least 4 pixel tiles per CTA
int pick_splits(int items, int total_tiles) {
  const int sms = ub_num_sms();
  int best = 1;
  double best_eff = 0.0;
  const int max_splits = total_tiles / 4 > 0 ? total_tiles / 4 : 1;
  for (int waves = 1; waves <= 4; ++waves) {
    int s = (waves * sms) / items;
    if (s < 1) s = 1;
    if (s > max_splits) s = max_splits;
    const long long ctas = (long long)items * s;
    const long long w = (ctas + sms - 1) / sms;
    // time ~ waves * (tiles per CTA + fixed prologue/epilogue cost of ~6 tile-times)
    const double tiles_per_cta = (double)total_tiles / s;
    const double cost = (double)w * (tiles_per_cta + 6.0);
    const double eff = 1.0 / cost;
    if (eff > best_eff * 1.02) {
      best_eff = eff;
      best = s;
    }
  }
  return best;
}

struct Plan {
  int mode, items, splits, n_tiles, tiles_w, tiles_h, total_tiles;
};

int make_plan(Plan& pl, int C0, int C1, int Cout, int N, int H, int W) {
  pl.mode = pick_mode(C0, C1, Cout);
  const int cblk = (C0 + C1) / 64;
  pl.tiles_w = (W + TW - 1) / TW;
  pl.tiles_h = (H + TH - 1) / TH;
  const long long tt = (long long)N * pl.tiles_w * pl.tiles_h;
  UB_CHECK_SHAPE(tt > 0 && tt < (1ll << 31), "wgrad: tile count out of range");
  pl.total_tiles = (int)tt;
  if (pl.mode == 0) {
    pl.n_tiles = Cout / 128;
    pl.items = (cblk / 2) * 3 * pl.n_tiles;
  } else {
    pl.n_tiles = Cout / 64;
    pl.items = cblk * pl.n_tiles;
  }
  pl.splits = pick_splits(pl.items, pl.total_tiles);
  return UB_OK;
}

template <int MODE>
int launch_wh(WgHaloParams& p, cudaStream_t stream) {
  using L = WhSmem<MODE>;
  static_assert(L::TOTAL <= 232448, "smem budget");
  auto kern = wgrad_halo_kernel<MODE>;
  static bool attr_done = false;
  if (!attr_done) {
    UB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_done = true;
  }
  kern<<<p.items * p.splits, 192, L::TOTAL, stream>>>(p);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

}  // namespace

long long ub_wgrad_halo_workspace_bytes(int C0, int C1, int Cout, int N, int H, int W) {
  Plan pl;
  if (make_plan(pl, C0, C1, Cout, N, H, W)) return -1;
  return pl.splits > 1 ? (long long)pl.splits * Cout * 9 * (C0 + C1) * 4 : 16;
}

int ub_wgrad_halo(const void* x0, int C0, const void* x1, int C1, const void* dz, int Cout, float* dw, void* workspace,
                  long long workspace_bytes, int N, int H, int W, cudaStream_t stream) {
  Plan pl;
  int rc = make_plan(pl, C0, C1, Cout, N, H, W);
  if (rc) return rc;
  const int Cin = C0 + C1;
  const long long n_w = (long long)Cout * 9 * Cin;
  WgHaloParams p;
  memset(&p, 0, sizeof(p));
  const int ph = pl.mode == 0 ? Cfg<0>::PH : Cfg<1>::PH;
  if ((rc = ub_tmap_act4d(&p.x_map[0], x0, C0, W, H, N, (long long)C0 * 2, (long long)W * C0 * 2, (long long)H * W * C0 * 2, PW, ph))) return rc;
  if (C1 > 0 && (rc = ub_tmap_act4d(&p.x_map[1], x1, C1, W, H, N, (long long)C1 * 2, (long long)W * C1 * 2, (long long)H * W * C1 * 2, PW, ph)))
    return rc;
  if ((rc = ub_tmap_act4d(&p.dz_map, dz, Cout, W, H, N, (long long)Cout * 2, (long long)W * Cout * 2, (long long)H * W * Cout * 2, TW, TH)))
    return rc;
  p.cblk0 = C0 / 64;
  p.cblk_total = Cin / 64;
  p.Cin = Cin;
  p.Cout = Cout;
  p.tiles_w = pl.tiles_w;
  p.tiles_h = pl.tiles_h;
  p.total_tiles = pl.total_tiles;
  p.items = pl.items;
  p.splits = pl.splits;
  p.n_tiles = pl.n_tiles;
  if (pl.splits > 1) {
    const long long need = (long long)pl.splits * n_w * 4;
    UB_CHECK_ARG(workspace && workspace_bytes >= need, "wgrad: workspace too small (%lld < %lld)", workspace_bytes, need);
    p.out = reinterpret_cast<float*>(workspace);
    p.split_stride = n_w;
  } else {
    p.out = dw;
    p.split_stride = 0;
  }
  static int p2 = -1;
  if (p2 < 0) {
    const char* e = getenv("UB_WGRAD_P_STAGES");      // 2: shallower pipeline, smaller shared-memory footprint (co-residency experiments)
    p2 = (e && e[0] == '2') ? 1 : 0;
  }
  rc = pl.mode == 0 ? (p2 ? launch_wh<2>(p, stream) : launch_wh<0>(p, stream)) : launch_wh<1>(p, stream);
  if (rc) return rc;
  if (pl.splits > 1) {
    const long long n4 = n_w / 4;      // n_w is a multiple of 64 * 64 * 9
    long long g = (n4 + 255) / 256;
    const long long cap = (long long)ub_num_sms() * 8;
    if (g > cap) g = cap;
    wgrad_sum_splits_kernel<<<(int)g, 256, 0, stream>>>(reinterpret_cast<const float4*>(workspace), reinterpret_cast<float4*>(dw), pl.splits, n4);
    UB_LAUNCH_CHECK();
  }
  return UB_OK;
}
