// conv3x3 ('same') forward / dgrad as an implicit GEMM on tcgen05/TMEM with ONE halo'd activation patch per 64-channel
// block: the nine filter taps are read by the tensor core as nine shifted views of that patch (sm_100a).
//
//   D[pixel, col] = sum_{tap, src, c}  X_src[pixel + shift(tap), c] * Wt[col, (tap, src, c)]        UNet/model.py:30-35, :57
//
// M tile = 128 pixels = 16 (h) x 8 (w) patch of one image.  TMA loads the {64 ch, 10 w, 18 h} halo'd box once
// (SWIZZLE_128B, out-of-bounds = zero = the 'same' padding); in that image pixel (pr, pc) is the 128-byte row
// pr * 10 + pc.  For tap (dh, dw) the A operand of tcgen05.mma is the K-major view starting at row
// (1 + dh) * 10 + (1 + dw) whose 8-row groups (= one tile row of 8 pixels) are 10 rows = 1280 bytes apart (SBO).  The
// start is off the 1024-byte swizzle atom; that is legal because the XOR pattern is a function of the absolute
// shared-memory address for both the TMA write and the MMA read (measured: tools/desc_probe.py, csrc/debug.cu).
// Compared with one TMA box per tap this cuts the activation bytes entering shared memory by 9 x 128 / 180 = 6.4x.
//
// Weights stream through a ring of per-(tap, channel-block) tiles; when the whole [BLOCK_N x 9 Cin] slice fits the
// ring (the 64-channel level-1 layers) it is loaded once per CTA and stays resident.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2-5 = epilogue (bias, ReLU, optional
// folded inference BatchNorm, bf16 store through TMA, per-channel sum / sum-of-squares partials for training BN).
#include "common.cuh"

namespace {

constexpr int TW = 8, TH = 16;
constexpr int PW = TW + 2, PH = TH + 2;
constexpr int PATCH_BYTES = PW * PH * 128;   // 23040
constexpr int PATCH_STRIDE = 23 * 1024;      // ring pitch (1024-aligned)
constexpr int OUT_BLK = 128 * 128;           // 128 pixels x 64 bf16

struct Conv3Params {
  CUtensorMap a_map[2];
  CUtensorMap b_map;
  CUtensorMap o_map[2];
  int nsrc;
  int cblk[2];
  int cblk_total;
  int H, W;
  int tiles_w, tiles_h;
  int n_tiles, total_tiles;
  int blocks_per_omap;
  const float* bias;         // nullable
  const float* post_scale;   // nullable: y = act(acc + bias) * post_scale + post_shift (inference BatchNorm folded)
  const float* post_shift;
  int relu;
  float* stats;              // [UB_STATS_ROWS][2][ncols] or null
  int ncols;
  int b_resident;
};

template <int BLOCK_N, int A_STAGES, int B_SLOTS>
struct C3Smem {
  static constexpr int B_BYTES = BLOCK_N * 128;
  static constexpr int OFF_B = A_STAGES * PATCH_STRIDE;
  static constexpr int OFF_OUT = OFF_B + B_SLOTS * B_BYTES;
  static constexpr int OUT_BYTES = (BLOCK_N / 64) * OUT_BLK;
  static constexpr int OFF_STAT = OFF_OUT + OUT_BYTES;              // float[4][2][BLOCK_N]
  static constexpr int OFF_VEC = OFF_STAT + 4 * 2 * BLOCK_N * 4;    // bias, scale, shift: float[3][BLOCK_N]
  static constexpr int OFF_BAR = OFF_VEC + 3 * BLOCK_N * 4;
  static constexpr int NBAR = 2 * A_STAGES + 2 * B_SLOTS + 4;
  static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
  static constexpr int TOTAL = OFF_TMEM + 16 + 1024;
};

__device__ __forceinline__ float c3_col_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

template <int BLOCK_N, int A_STAGES, int B_SLOTS>
__global__ void __launch_bounds__(192, 1) conv3_kernel(const __grid_constant__ Conv3Params p) {
  using L = C3Smem<BLOCK_N, A_STAGES, B_SLOTS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* out_stage = smem + L::OFF_OUT;
  float* stat_smem = reinterpret_cast<float*>(smem + L::OFF_STAT);
  float* vec_smem = reinterpret_cast<float*>(smem + L::OFF_VEC);
  uint64_t* afull = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* aempty = afull + A_STAGES;
  uint64_t* bfull = aempty + A_STAGES;
  uint64_t* bempty = bfull + B_SLOTS;
  uint64_t* tfull = bempty + B_SLOTS;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::OFF_TMEM);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.nsrc; ++i) tma_prefetch_desc(&p.a_map[i]);
    tma_prefetch_desc(&p.b_map);
    tma_prefetch_desc(&p.o_map[0]);
    for (int s = 0; s < A_STAGES; ++s) {
      mbar_init(&afull[s], 1);
      mbar_init(&aempty[s], 1);
    }
    for (int s = 0; s < B_SLOTS; ++s) {
      mbar_init(&bfull[s], 1);
      mbar_init(&bempty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);
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
  const int n_tile = blockIdx.x % p.n_tiles;       // host guarantees gridDim.x % n_tiles == 0: fixed per CTA
  const bool resident = p.b_resident != 0;

  if (warp == 0) {
    // ================= TMA producer (all lanes run the loop, one elected lane issues) =================
    int a_stage = 0, b_slot = 0;
    uint32_t a_phase = 0, b_phase = 0;
    bool first = true;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_tiles;
      const int img = m_tile / tiles_per_img;
      const int rem = m_tile - img * tiles_per_img;
      const int h0 = (rem / p.tiles_w) * TH;
      const int w0 = (rem % p.tiles_w) * TW;
      int cbg = 0;
      for (int src = 0; src < p.nsrc; ++src) {
        for (int cb = 0; cb < p.cblk[src]; ++cb, ++cbg) {
          mbar_wait(&aempty[a_stage], a_phase ^ 1);
          mbar_expect_tx_e(&afull[a_stage], PATCH_BYTES);
          tma_load_4d_e(smem + a_stage * PATCH_STRIDE, &p.a_map[src], &afull[a_stage], cb * 64, w0 - 1, h0 - 1, img);
          if (++a_stage == A_STAGES) { a_stage = 0; a_phase ^= 1; }
          if (resident && !first) continue;
          for (int tap = 0; tap < 9; ++tap) {
            const int slot = resident ? cbg * 9 + tap : b_slot;
            if (!resident) mbar_wait(&bempty[slot], b_phase ^ 1);
            mbar_expect_tx_e(&bfull[slot], L::B_BYTES);
            tma_load_2d_e(smem + L::OFF_B + slot * L::B_BYTES, &p.b_map, &bfull[slot], (tap * p.cblk_total + cbg) * 64, n_tile * BLOCK_N);
            if (!resident && ++b_slot == B_SLOTS) { b_slot = 0; b_phase ^= 1; }
          }
        }
      }
      first = false;
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer (all lanes run the loop, one elected lane issues) =================
    constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 0, 0);
    const uint32_t tmem_u = warp_uniform(tmem_base);
    const uint32_t smem_base_u = warp_uniform(smem_u32(smem));
    int a_stage = 0, b_slot = 0;
    uint32_t a_phase = 0, b_phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_u + as * BLOCK_N;
      for (int cbg = 0; cbg < p.cblk_total; ++cbg) {
        mbar_wait(&afull[a_stage], a_phase);
        const uint32_t patch = smem_base_u + a_stage * PATCH_STRIDE;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int slot = resident ? cbg * 9 + tap : b_slot;
          mbar_wait(&bfull[slot], resident ? 0u : b_phase);
          tc_fence_after();
          const int dh = tap / 3, dw = tap % 3;           // already offset by +1 (patch origin is pixel (-1, -1))
          const uint64_t adesc = make_smem_desc(patch + (dh * PW + dw) * 128, 16, PW * 128);
          const uint64_t bdesc = make_smem_desc(smem_base_u + L::OFF_B + slot * L::B_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_bf16_e(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (cbg | tap | k) != 0);
          if (!resident) {
            tc_commit_e(&bempty[slot]);
            if (++b_slot == B_SLOTS) { b_slot = 0; b_phase ^= 1; }
          }
        }
        tc_commit_e(&aempty[a_stage]);
        if (++a_stage == A_STAGES) { a_stage = 0; a_phase ^= 1; }
      }
      tc_commit_e(&tfull[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    __syncwarp();
  } else {
    // ================= epilogue (4 warps, one TMEM lane quadrant each) =================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int et = threadIdx.x - 64;          // 0..127
    const bool store_thread = (et == 0);
    const uint32_t row_smem = smem_u32(out_stage) + row * 128;
    const int rsw = row & 7;
    constexpr int NCHUNK = BLOCK_N / 32;
    float acc_sum[NCHUNK], acc_sq[NCHUNK];
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) acc_sum[c] = acc_sq[c] = 0.f;
    const bool post = p.post_scale != nullptr;
    for (int c = et; c < BLOCK_N; c += 128) {
      const int col = n_tile * BLOCK_N + c;
      vec_smem[c] = p.bias ? p.bias[col] : 0.f;
      vec_smem[BLOCK_N + c] = post ? p.post_scale[col] : 1.f;
      vec_smem[2 * BLOCK_N + c] = post ? p.post_shift[col] : 0.f;
    }
    named_bar_sync(1, 128);
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_tiles;
      const int img = m_tile / tiles_per_img;
      const int rem = m_tile - img * tiles_per_img;
      const int h0 = (rem / p.tiles_w) * TH;
      const int w0 = (rem % p.tiles_w) * TW;
      const bool valid = (h0 + row / TW < p.H) && (w0 + row % TW < p.W);

      // staging buffer must have been drained by the previous tile's TMA store
      if (store_thread) tma_store_wait_read0();
      named_bar_sync(1, 128);

      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
#pragma unroll
      for (int chunk = 0; chunk < NCHUNK; ++chunk) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + as * BLOCK_N + chunk * 32, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(v[j]) + vec_smem[chunk * 32 + j];
          f[j] = p.relu ? fmaxf(x, 0.f) : x;
        }
        if (post) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaf(f[j], vec_smem[BLOCK_N + chunk * 32 + j], vec_smem[2 * BLOCK_N + chunk * 32 + j]);
        }
        const uint32_t blk = row_smem + (chunk >> 1) * OUT_BLK;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c16 = (chunk & 1) * 4 + q;
          st_shared_v4(blk + ((c16 ^ rsw) << 4), pack_bf16x2(f[q * 8 + 0], f[q * 8 + 1]), pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]),
                       pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]), pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]));
        }
        if (p.stats) {
          float s[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            f[j] = valid ? f[j] : 0.f;
            s[j] = f[j] * f[j];
          }
          acc_sum[chunk] += c3_col_reduce32(f, lane);
          acc_sq[chunk] += c3_col_reduce32(s, lane);
        }
      }
      // accumulator stage drained -> MMA warp may reuse it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      // staged tile -> global through TMA (clipped at the image edge)
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (store_thread) {
#pragma unroll
        for (int b = 0; b < BLOCK_N / 64; ++b) {
          const int j = n_tile * (BLOCK_N / 64) + b;
          const int map = j / p.blocks_per_omap;
          const int c0 = (j - map * p.blocks_per_omap) * 64;
          tma_store_4d(&p.o_map[map], out_stage + b * OUT_BLK, c0, w0, h0, img);
        }
        tma_store_commit();
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if (store_thread) tma_store_wait_all0();
    if (p.stats) {
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        stat_smem[(quad * 2 + 0) * BLOCK_N + c * 32 + lane] = acc_sum[c];
        stat_smem[(quad * 2 + 1) * BLOCK_N + c * 32 + lane] = acc_sq[c];
      }
      named_bar_sync(1, 128);
      const int srow = blockIdx.x / p.n_tiles;
      for (int i = et; i < 2 * BLOCK_N; i += 128) {
        const int which = i / BLOCK_N, c = i % BLOCK_N;
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) t += stat_smem[(q * 2 + which) * BLOCK_N + c];
        p.stats[((size_t)srow * 2 + which) * p.ncols + n_tile * BLOCK_N + c] = t;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BLOCK_N);
  }
}

template <int BLOCK_N, int A_STAGES, int B_SLOTS>
int launch_c3(Conv3Params& p, int n_img, cudaStream_t stream) {
  using L = C3Smem<BLOCK_N, A_STAGES, B_SLOTS>;
  static_assert(L::TOTAL <= 232448, "smem budget");
  auto kern = conv3_kernel<BLOCK_N, A_STAGES, B_SLOTS>;
  static bool attr_done = false;
  if (!attr_done) {
    UB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_done = true;
  }
  p.n_tiles = p.ncols / BLOCK_N;
  p.tiles_w = (p.W + TW - 1) / TW;
  p.tiles_h = (p.H + TH - 1) / TH;
  const long long m_tiles = (long long)n_img * p.tiles_w * p.tiles_h;
  const long long total = m_tiles * p.n_tiles;
  UB_CHECK_SHAPE(total > 0 && total < (1ll << 31), "conv3: tile count out of range");
  p.total_tiles = (int)total;
  p.b_resident = (9 * p.cblk_total <= B_SLOTS) ? 1 : 0;
  const int sms = ub_num_sms();
  long long grid = (long long)(sms / p.n_tiles) * p.n_tiles;
  if (grid <= 0) grid = p.n_tiles;
  if (grid > total) grid = total;            // total is a multiple of n_tiles
  UB_CHECK_SHAPE(grid / p.n_tiles <= UB_STATS_ROWS, "conv3: stats rows");
  if (p.stats) UB_CUDA(cudaMemsetAsync(p.stats, 0, sizeof(float) * UB_STATS_ROWS * 2 * p.ncols, stream));
  kern<<<(int)grid, 192, L::TOTAL, stream>>>(p);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

int launch(Conv3Params& p, int n_img, cudaStream_t stream) {
  p.cblk_total = 0;
  for (int i = 0; i < p.nsrc; ++i) p.cblk_total += p.cblk[i];
  UB_CHECK_SHAPE(p.ncols % 64 == 0 && p.cblk_total > 0, "conv3: columns must be a multiple of 64");
  if (p.ncols % 256 == 0) return launch_c3<256, 2, 3>(p, n_img, stream);
  if (p.ncols % 128 == 0) return launch_c3<128, 3, 6>(p, n_img, stream);
  return launch_c3<64, 2, 18>(p, n_img, stream);
}

int in_map(CUtensorMap* m, const void* base, int C, int W, int H, int N) {
  return ub_tmap_act4d(m, base, C, W, H, N, (long long)C * 2, (long long)W * C * 2, (long long)H * W * C * 2, PW, PH);
}
int out_map(CUtensorMap* m, const void* base, int C, int W, int H, int N) {
  return ub_tmap_act4d(m, base, C, W, H, N, (long long)C * 2, (long long)W * C * 2, (long long)H * W * C * 2, TW, TH);
}

}  // namespace

// Called by the extern "C" entry points in igemm_fwd.cu.
int ub_conv3_halo_fwd(const void* x0, int C0, const void* x1, int C1, const void* w, const float* bias, const float* post_scale,
                      const float* post_shift, void* out, float* stats, int N, int H, int W, int Cout, int relu, cudaStream_t stream) {
  Conv3Params p;
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = in_map(&p.a_map[0], x0, C0, W, H, N))) return rc;
  p.nsrc = 1;
  p.cblk[0] = C0 / 64;
  if (C1 > 0) {
    if ((rc = in_map(&p.a_map[1], x1, C1, W, H, N))) return rc;
    p.nsrc = 2;
    p.cblk[1] = C1 / 64;
  }
  const int Cin = C0 + C1;
  const int bn = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0 ? 128 : 64);
  if ((rc = ub_tmap_mat2d(&p.b_map, w, Cout, 9ll * Cin, bn))) return rc;
  if ((rc = out_map(&p.o_map[0], out, Cout, W, H, N))) return rc;
  p.H = H;
  p.W = W;
  p.blocks_per_omap = Cout / 64;
  p.bias = bias;
  p.post_scale = post_scale;
  p.post_shift = post_shift;
  p.relu = relu;
  p.stats = stats;
  p.ncols = Cout;
  return launch(p, N, stream);
}

int ub_conv3_halo_dgrad(const void* dz, int Cout, const void* w_t, void* dx0, int C0, void* dx1, int C1, int N, int H, int W,
                        cudaStream_t stream) {
  Conv3Params p;
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = in_map(&p.a_map[0], dz, Cout, W, H, N))) return rc;
  p.nsrc = 1;
  p.cblk[0] = Cout / 64;
  const int Cin = C0 + C1;
  const int bn = (Cin % 256 == 0) ? 256 : (Cin % 128 == 0 ? 128 : 64);
  if ((rc = ub_tmap_mat2d(&p.b_map, w_t, Cin, 9ll * Cout, bn))) return rc;
  if ((rc = out_map(&p.o_map[0], dx0, C0, W, H, N))) return rc;
  if (C1 > 0 && (rc = out_map(&p.o_map[1], dx1, C1, W, H, N))) return rc;
  p.H = H;
  p.W = W;
  p.blocks_per_omap = C0 / 64;
  p.relu = 0;
  p.ncols = Cin;
  return launch(p, N, stream);
}
