// Shared epilogue of the implicit-GEMM kernels (sm_100a): 8 warps (two per TMEM lane quadrant, each taking half of the
// tile's columns) drain a 128 x BLOCK_N fp32 accumulator stage:
//   acc + bias -> ReLU -> optional folded inference BatchNorm -> bf16 -> SWIZZLE_128B staging tile -> TMA store,
// and, for training, the per-channel sum / sum-of-squares of the STORED (bf16-rounded) activations for the BatchNorm
// that follows (UNet/model.py:36, :47).  The statistics are column sums of the staging tile, read back with a
// column-pair-per-thread mapping (conflict-free 4-byte reads, 4 accumulators per thread for the whole kernel) -- the
// first version reduced the fp32 accumulators across lanes with a 31-shuffle butterfly per 32 columns, which made the
// 64/128-column tiles epilogue-bound (profiles/r01_*).
#pragma once
#include "common.cuh"

constexpr int EPI_WARPS = 8;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int EPI_OUT_BLK = 128 * 128;   // 128 pixels x 64 bf16

struct EpiParams {
  const float* bias;         // nullable, indexed (column % bias_mod)
  int bias_mod;
  const float* post_scale;   // nullable, same indexing: y = act(acc + bias) * post_scale + post_shift
  const float* post_shift;
  int relu;
  float* stats;              // [UB_STATS_ROWS][2][ncols] or null
  int ncols;
  int H, W;                  // pixel space of the tile grid (for masking ragged tiles out of the statistics)
  // RED (dgrad only): the tile being written is dL/dy of a BatchNorm'd tensor whose saved activation `a` is read through
  // red_map; the epilogue accumulates that BatchNorm's backward sums (the separate bn_bwd_reduce pass over dy and a):
  //   red_out[row][0][c] = sum dy,  red_out[row][1][c] = rstd_c * sum dy * (a - mean_c)        (c relative to red_blk_begin * 64)
  const float* red_mean;
  const float* red_rstd;
  float* red_out;            // [UB_STATS_ROWS][2][red_ncols]
  int red_blk_begin;         // first 64-column block of the output that belongs to the BatchNorm'd tensor (concat dgrad: C0 / 64)
  int red_ncols;
  // FIN (forward with statistics): the LAST CTA to finish sums the partial rows of every column in a fixed order (fp64) and
  // finalises the BatchNorm that follows -- mean, rstd and the moving statistics -- so that no separate ub_bn_finalize launch (and
  // no zero-fill of unused partial rows) sits between this kernel and its consumer.  fin_counter: zero-initialised, reset by the
  // last CTA.  Channel c gathers the columns g * (ncols / fin_groups) + c (deconv: 4 column groups per channel).
  float* fin_mean;           // null = off
  float* fin_rstd;
  float* fin_moving_mean;    // nullable
  float* fin_moving_var;
  unsigned int* fin_counter;
  double fin_count;          // samples per channel (N * H * W of the BatchNorm'd tensor)
  float fin_momentum, fin_eps;
  int fin_groups;
  int fin_rows;              // partial rows written by this launch (gridDim.x / n_tiles)
};

// host-side description of the BatchNorm a forward launch finalises itself (EpiParams::fin_*)
struct BnFin {
  float* mean;
  float* rstd;
  float* moving_mean;
  float* moving_var;
  double count;
  float momentum, eps;
  int groups;
  unsigned int* counter;
};
inline void epi_set_fin(EpiParams& ep, const BnFin& f, int rows) {
  ep.fin_mean = f.mean;
  ep.fin_rstd = f.rstd;
  ep.fin_moving_mean = f.moving_mean;
  ep.fin_moving_var = f.moving_var;
  ep.fin_counter = f.counter;
  ep.fin_count = f.count;
  ep.fin_momentum = f.momentum;
  ep.fin_eps = f.eps;
  ep.fin_groups = f.groups;
  ep.fin_rows = rows;
}

// bytes of epilogue scratch the fused finalise needs: 256 x 8 doubles of row-lane partials + 2 x ncols doubles of column totals
__host__ __device__ constexpr int epi_fin_bytes(int ncols) { return 256 * 8 * 8 + 2 * ncols * 8; }

// RED: 0 = off, 1 = one `a`-tile buffer (fetched at the start of its part), 2 = two buffers (fetched one part ahead)
// CASEB: the bias is a [9][ncols] table indexed by the pixel's border case (3 row cases x 3 column cases) -- the forward
// convolution of a BatchNorm-FOLDED input (`ub_conv3x3_fwd_cases`): 'same' padding is applied after BatchNorm, so the folded
// shift contributes only through the taps that lie inside the image (oracle/unet_numpy.py border_case_bias)
template <int BLOCK_N, int OUT_BUFS, int RED = 0, int CASEB = 0>
struct EpiSmem {
  static constexpr int OUT_BYTES = (BLOCK_N / 64) * EPI_OUT_BLK;
  static constexpr int OFF_STAT = 0;                                  // float[row groups][2][BLOCK_N] (16 KB): aliases staging buffer 0,
                                                                      // only touched in finish() after every TMA store has drained
  static constexpr int OFF_ABUF = OUT_BUFS * OUT_BYTES;               // RED: the `a` tile, same swizzled layout as the staging tile
  static constexpr int OFF_VEC = OFF_ABUF + RED * OUT_BYTES;          // scale, shift, bias: float[2 + NBIAS][BLOCK_N]
  static constexpr int NBIAS = CASEB ? 9 : 1;
  static constexpr int OFF_ABAR = OFF_VEC + (2 + NBIAS) * BLOCK_N * 4;   // RED: mbarrier of the `a` tile load
  static constexpr int TOTAL = OFF_ABAR + 16;
  static_assert((256 / (BLOCK_N / 8)) * 2 * BLOCK_N * 4 <= OUT_BYTES, "statistics scratch must fit the staging buffer");
};

// TILE_W_: pixels per tile row (tile row r of accumulator row m: m / TILE_W_, column m % TILE_W_)
template <int BLOCK_N, int OUT_BUFS, int TILE_W_, int RED = 0, int CASEB = 0>
struct Epilogue {
  using S = EpiSmem<BLOCK_N, OUT_BUFS, RED, CASEB>;
  static constexpr int COLS_PER_THREAD = BLOCK_N / 2;          // this warp's half of the columns
  static constexpr int NCHUNK = COLS_PER_THREAD / 32;
  static constexpr int PAIRS = BLOCK_N / 2;                    // column pairs
  static constexpr int ROW_GROUPS = EPI_THREADS / PAIRS;       // 8 / 4 / 2 for BLOCK_N 64 / 128 / 256
  static constexpr int ROWS_PER_GROUP = 128 / ROW_GROUPS;
  // forward statistics: thread = (16-byte chunk of 8 columns, row group); rows are interleaved over the groups so that the four rows a
  // warp reads at once differ in (row & 7) -- with the 128-byte swizzle every ld.shared.v4 wavefront then covers all 32 banks
  static constexpr int ST_CHUNKS = BLOCK_N / 8;                // 8 / 16 / 32
  static constexpr int ST_GROUPS = EPI_THREADS / ST_CHUNKS;    // 32 / 16 / 8
  static constexpr int ST_ROWS = 128 / ST_GROUPS;              // rows per thread and part: 4 / 8 / 16

  uint8_t* base;             // epilogue smem region (1024-aligned)
  const EpiParams& ep;
  uint32_t tmem_base;
  uint64_t *tfull, *tempty;
  int et, quad, lane, half, row;
  int as = 0, buf = 0;
  uint32_t aphase = 0;
  bool post;
  // forward statistics: sum / sum of squares of this thread's 8 columns; RED: sum dy / sum dy * (a - mean) of the same 8 columns
  float ss[8] = {}, qq[8] = {};
  // RED state
  const CUtensorMap* red_map = nullptr;
  int red_nt = 0;            // this CTA's n_tile
  uint32_t red_phase0 = 0, red_phase1 = 0;
  int red_cur = 0;           // buffer holding the current part's `a` tile
  bool red_primed = false;
  bool red_mine = false;     // this thread's 8 columns belong to the BatchNorm'd tensor
  int pair_rank = -1;        // >= 0: this CTA is rank `pair_rank` of a cta_group::2 pair (igemm_conv3_2cta.cu)

  // epi_thread: 0 .. EPI_THREADS-1; hw_warp: warp index within the CTA (a warp may only touch TMEM lanes 32 * (hw_warp % 4) .. +31)
  __device__ __forceinline__ Epilogue(uint8_t* smem_region, const EpiParams& e, uint32_t tmem, uint64_t* tf, uint64_t* te, int epi_thread,
                                      int hw_warp)
      : base(smem_region), ep(e), tmem_base(tmem), tfull(tf), tempty(te), et(epi_thread) {
    lane = et & 31;
    quad = hw_warp & 3;
    half = (et >> 5) >> 2;     // the two warps sharing a quadrant split the columns
    row = quad * 32 + lane;
    post = ep.post_scale != nullptr;
  }

  __device__ __forceinline__ float* vec() const { return reinterpret_cast<float*>(base + S::OFF_VEC); }
  static __device__ __forceinline__ int bias_off(int k) { return k ? (2 + k) * BLOCK_N : 0; }

  // column-dependent vectors of this CTA's n_tile (call once, or when n_tile changes); ends with a barrier
  __device__ __forceinline__ void load_vectors(int n_tile) {
    float* v = vec();
    for (int c = et; c < BLOCK_N; c += EPI_THREADS) {
      const int col = (n_tile * BLOCK_N + c) % ep.bias_mod;
      v[c] = ep.bias ? ep.bias[col] : 0.f;
      v[BLOCK_N + c] = post ? ep.post_scale[col] : 1.f;
      v[2 * BLOCK_N + c] = post ? ep.post_shift[col] : 0.f;
      if (CASEB) {          // ep.bias = [9][ncols]: case 0 sits in the plain bias slot, cases 1..8 behind scale / shift
#pragma unroll
        for (int k = 1; k < 9; ++k) v[bias_off(k) + c] = ep.bias[(size_t)k * ep.ncols + col];
      }
    }
    if (RED) {
      red_nt = n_tile;
      const int c = 8 * (et % ST_CHUNKS);
      const int gcol = n_tile * BLOCK_N + c - ep.red_blk_begin * 64;
      red_mine = gcol >= 0 && gcol < ep.red_ncols;          // whole 64-column blocks belong to the tensor or do not
      // the column means sit in the (otherwise unused: a dgrad has no bias) bias slots of the vector area
      for (int cc = et; cc < BLOCK_N; cc += EPI_THREADS) {
        const int g = n_tile * BLOCK_N + cc - ep.red_blk_begin * 64;
        v[cc] = (g >= 0 && g < ep.red_ncols) ? ep.red_mean[g] : 0.f;
      }
      if (et == 0) {
        mbar_init(abar(0), 1);
        mbar_init(abar(1), 1);
        mbar_fence_init();
      }
    }
    named_bar_sync(1, EPI_THREADS);
  }
  __device__ __forceinline__ uint64_t* abar(int i) const { return reinterpret_cast<uint64_t*>(base + S::OFF_ABAR) + i; }
  __device__ __forceinline__ bool red_any() const {
    bool any = false;
#pragma unroll
    for (int b = 0; b < BLOCK_N / 64; ++b) any |= red_block(b);
    return any;
  }
  // one thread: fetch the tile of the saved activation matching the output part at (h0, w0, img) -- same box and swizzle as
  // the staging tile -- into buffer `bi`
  __device__ __forceinline__ void red_issue(int bi, int h0, int w0, int img) {
    uint32_t bytes = 0;
#pragma unroll
    for (int b = 0; b < BLOCK_N / 64; ++b) bytes += red_block(b) ? EPI_OUT_BLK : 0;
    if (!bytes) return;
    mbar_expect_tx(abar(bi), bytes);
#pragma unroll
    for (int b = 0; b < BLOCK_N / 64; ++b)
      if (red_block(b))
        tma_load_4d(base + S::OFF_ABUF + bi * S::OUT_BYTES + b * EPI_OUT_BLK, red_map, abar(bi),
                    (red_nt * (BLOCK_N / 64) + b - ep.red_blk_begin) * 64, w0, h0, img);
  }
  __device__ __forceinline__ bool red_block(int b) const {
    const int jb = red_nt * (BLOCK_N / 64) + b - ep.red_blk_begin;
    return jb >= 0 && jb * 64 < ep.red_ncols;
  }

  // store_fn(staging_block_ptr, block_index_within_tile) issues the TMA store(s) of one 64-column block (one thread calls it)
  // One 128-pixel x BLOCK_N part of an accumulator stage.  A stage may hold several parts (the 256-pixel super-tiles of
  // igemm_conv3.cu): `col_off` = the part's first TMEM column inside the stage, `stage_cols` = columns per stage,
  // `first` waits for the stage's MMAs, `last` hands the stage back to the MMA warp.
  template <typename StoreFn>
  __device__ __forceinline__ void tile(int h0, int w0, StoreFn&& store_fn, int col_off = 0, bool first = true, bool last = true,
                                       int stage_cols = BLOCK_N, int img = 0, bool has_next = false, int nh0 = 0, int nw0 = 0, int nimg = 0) {
    uint8_t* out_stage = base + buf * S::OUT_BYTES;
    // the staging buffer we are about to overwrite must have been read by its TMA store (OUT_BUFS - 1 stores may be in flight)
    if (et == 0) {
      if (OUT_BUFS == 1) tma_store_wait_read0();
      else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    }
    named_bar_sync(1, EPI_THREADS);      // also orders the previous tile's column-sum reads before these writes
    if (RED) {
      // the buffer not holding the current part was last read one part ago (ordered by the barrier above)
      if (et == 0) {
        if (RED == 1 || !red_primed) red_issue(red_cur, h0, w0, img);
        if (RED == 2 && has_next) red_issue(red_cur ^ 1, nh0, nw0, nimg);
      }
      red_primed = true;
    }

    if (first) mbar_wait(&tfull[as], aphase);
    tc_fence_after();
    const float* v = vec();
    const float* vb = v;
    if (CASEB) {
      // border case of this thread's pixel: row case * 3 + column case, 0 = first, 1 = interior, 2 = last (H, W >= 2)
      const int hh = h0 + row / TILE_W_, ww = w0 + row % TILE_W_;
      const int rc = hh <= 0 ? 0 : (hh >= ep.H - 1 ? 2 : 1), cc = ww <= 0 ? 0 : (ww >= ep.W - 1 ? 2 : 1);
      vb = v + bias_off(rc * 3 + cc);
    }
    const uint32_t row_smem = smem_u32(out_stage) + row * 128;
    const int rsw = row & 7;
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
      const int chunk = half * NCHUNK + ch;               // 32-column chunk index within the tile
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + as * stage_cols + col_off + chunk * 32, r);
      tmem_ld_wait();
      float f[32];
      const float4* vb4 = reinterpret_cast<const float4*>(vb + chunk * 32);     // 16-byte aligned: OFF_VEC and every vector start are
      if (RED != 0) {               // a dgrad: no bias, no activation (ub_conv3_halo_dgrad leaves both unset)
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(r[j]);
      } else
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b4 = vb4[j];
        const float x0 = __uint_as_float(r[4 * j + 0]) + b4.x, x1 = __uint_as_float(r[4 * j + 1]) + b4.y;
        const float x2 = __uint_as_float(r[4 * j + 2]) + b4.z, x3 = __uint_as_float(r[4 * j + 3]) + b4.w;
        f[4 * j + 0] = ep.relu ? fmaxf(x0, 0.f) : x0;
        f[4 * j + 1] = ep.relu ? fmaxf(x1, 0.f) : x1;
        f[4 * j + 2] = ep.relu ? fmaxf(x2, 0.f) : x2;
        f[4 * j + 3] = ep.relu ? fmaxf(x3, 0.f) : x3;
      }
      if (post) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = fmaf(f[j], v[BLOCK_N + chunk * 32 + j], v[2 * BLOCK_N + chunk * 32 + j]);
      }
      const uint32_t blk = row_smem + (chunk >> 1) * EPI_OUT_BLK;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c16 = (chunk & 1) * 4 + q;
        st_shared_v4(blk + ((c16 ^ rsw) << 4), pack_bf16x2(f[q * 8 + 0], f[q * 8 + 1]), pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]),
                     pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]), pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]));
      }
    }
    // accumulator stage drained -> the MMA warp may reuse it
    tc_fence_before();
    __syncwarp();
    if (last && lane == 0) {
      if (pair_rank < 0) mbar_arrive(&tempty[as]);
      else mbar_arrive_cluster(&tempty[as], 0);         // CTA pair: the MMA issuer (leader) waits for the epilogues of BOTH CTAs
    }
    fence_proxy_async_smem();
    named_bar_sync(1, EPI_THREADS);
    if (et == 0) {
#pragma unroll
      for (int b = 0; b < BLOCK_N / 64; ++b) store_fn(out_stage + b * EPI_OUT_BLK, b);
      tma_store_commit();
    }
    if (RED == 0 && ep.stats) {   // (a dgrad with a fused reduction never has forward statistics)
      // column sums of the staged (rounded) tile: thread = (16-byte chunk, row group); four independent 16-byte loads in flight
      const int ck = et % ST_CHUNKS, grp = et / ST_CHUNKS;
      const uint32_t cbase = smem_u32(out_stage) + (ck >> 3) * EPI_OUT_BLK;
      const int c16 = ck & 7;
#pragma unroll
      for (int rb = 0; rb < ST_ROWS; rb += 4) {
        uint32_t u[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r_ = (rb + i) * ST_GROUPS + grp;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(u[i][0]), "=r"(u[i][1]), "=r"(u[i][2]), "=r"(u[i][3])
                       : "r"(cbase + r_ * 128 + ((c16 ^ (r_ & 7)) << 4)));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r_ = (rb + i) * ST_GROUPS + grp;
          if ((h0 + r_ / TILE_W_ < ep.H) && (w0 + r_ % TILE_W_ < ep.W)) {      // ragged tiles: rows outside the image do not count
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float a = __uint_as_float(u[i][k] << 16), b = __uint_as_float(u[i][k] & 0xffff0000u);
              ss[2 * k] += a;
              ss[2 * k + 1] += b;
              qq[2 * k] = fmaf(a, a, qq[2 * k]);
              qq[2 * k + 1] = fmaf(b, b, qq[2 * k + 1]);
            }
          }
        }
      }
    }
    if (RED) {
      if (red_any()) {                             // uniform over the CTA
        if (red_cur == 0) {
          mbar_wait(abar(0), red_phase0);
          red_phase0 ^= 1;
        } else {
          mbar_wait(abar(1), red_phase1);
          red_phase1 ^= 1;
        }
        if (red_mine) {
          // thread = (16-byte chunk, row group) as for the forward statistics; dy from the staging tile, `a` from its TMA'd tile
          const int ck = et % ST_CHUNKS, grp = et / ST_CHUNKS;
          const uint32_t off0 = (ck >> 3) * EPI_OUT_BLK;
          const uint32_t dybase = smem_u32(out_stage) + off0;
          const uint32_t abase = smem_u32(base + S::OFF_ABUF + red_cur * S::OUT_BYTES) + off0;
          const int c16 = ck & 7;
          const float4 mlo = *reinterpret_cast<const float4*>(vec() + ck * 8), mhi = *reinterpret_cast<const float4*>(vec() + ck * 8 + 4);
          const float rm[8] = {mlo.x, mlo.y, mlo.z, mlo.w, mhi.x, mhi.y, mhi.z, mhi.w};
#pragma unroll
          for (int rb = 0; rb < ST_ROWS; rb += 2) {
            uint32_t u[2][4], v[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int r_ = (rb + i) * ST_GROUPS + grp;
              const uint32_t o = r_ * 128 + ((c16 ^ (r_ & 7)) << 4);
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u[i][0]), "=r"(u[i][1]), "=r"(u[i][2]), "=r"(u[i][3]) : "r"(dybase + o));
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[i][0]), "=r"(v[i][1]), "=r"(v[i][2]), "=r"(v[i][3]) : "r"(abase + o));
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int r_ = (rb + i) * ST_GROUPS + grp;
              if ((h0 + r_ / TILE_W_ < ep.H) && (w0 + r_ % TILE_W_ < ep.W)) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float d0 = __uint_as_float(u[i][k] << 16), d1 = __uint_as_float(u[i][k] & 0xffff0000u);
                  const float a0 = __uint_as_float(v[i][k] << 16), a1 = __uint_as_float(v[i][k] & 0xffff0000u);
                  ss[2 * k] += d0;
                  ss[2 * k + 1] += d1;
                  qq[2 * k] = fmaf(d0, a0 - rm[2 * k], qq[2 * k]);
                  qq[2 * k + 1] = fmaf(d1, a1 - rm[2 * k + 1], qq[2 * k + 1]);
                }
              }
            }
          }
        }
      }
      if (RED == 2) red_cur ^= 1;
    }
    if (last) {
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if (OUT_BUFS > 1) buf = (buf + 1 == OUT_BUFS) ? 0 : buf + 1;
  }

  // after the last tile: drain stores, write this CTA's partial statistics row
  __device__ __forceinline__ void finish(int n_tile, int stats_row) {
    if (et == 0) tma_store_wait_all0();
    if (RED) {
      float* st = reinterpret_cast<float*>(base + S::OFF_STAT);     // [ST_GROUPS][2][BLOCK_N]
      const int ck = et % ST_CHUNKS, grp = et / ST_CHUNKS;
      named_bar_sync(1, EPI_THREADS);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        st[(grp * 2 + 0) * BLOCK_N + ck * 8 + k] = ss[k];
        st[(grp * 2 + 1) * BLOCK_N + ck * 8 + k] = qq[k];
      }
      named_bar_sync(1, EPI_THREADS);
      for (int i = et; i < 2 * BLOCK_N; i += EPI_THREADS) {
        const int which = i / BLOCK_N, c = i % BLOCK_N;
        const int gcol = n_tile * BLOCK_N + c - ep.red_blk_begin * 64;
        if (gcol < 0 || gcol >= ep.red_ncols) continue;
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < ST_GROUPS; ++g) t += st[(g * 2 + which) * BLOCK_N + c];
        if (which) t *= ep.red_rstd[gcol];
        ep.red_out[((size_t)stats_row * 2 + which) * ep.red_ncols + gcol] = t;
      }
      return;
    }
    if (ep.stats) {
      float* st = reinterpret_cast<float*>(base + S::OFF_STAT);     // [ST_GROUPS][2][BLOCK_N]
      const int ck = et % ST_CHUNKS, grp = et / ST_CHUNKS;
      named_bar_sync(1, EPI_THREADS);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        st[(grp * 2 + 0) * BLOCK_N + ck * 8 + k] = ss[k];
        st[(grp * 2 + 1) * BLOCK_N + ck * 8 + k] = qq[k];
      }
      named_bar_sync(1, EPI_THREADS);
      for (int i = et; i < 2 * BLOCK_N; i += EPI_THREADS) {
        const int which = i / BLOCK_N, c = i % BLOCK_N;
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < ST_GROUPS; ++g) t += st[(g * 2 + which) * BLOCK_N + c];
        ep.stats[((size_t)stats_row * 2 + which) * ep.ncols + n_tile * BLOCK_N + c] = t;
      }
      if (ep.fin_mean) finalize_last_cta(st);
    }
  }

  // see EpiParams::fin_*.  Called by all epilogue threads after this CTA's partial row has been written.
  __device__ __forceinline__ void finalize_last_cta(float* scratch) {
    __threadfence();                                   // this thread's partial-row stores are visible device-wide ...
    named_bar_sync(1, EPI_THREADS);
    volatile unsigned int* slot = reinterpret_cast<volatile unsigned int*>(scratch);
    if (et == 0) *slot = atomicAdd(ep.fin_counter, 1u);   // ... before this CTA is counted as done
    named_bar_sync(1, EPI_THREADS);
    const unsigned int ticket = *slot;
    if (ticket != gridDim.x - 1) return;               // uniform over the CTA
    __threadfence();
    named_bar_sync(1, EPI_THREADS);                    // everyone has read the ticket before the scratch is reused
    const int ncols = ep.ncols, rows = ep.fin_rows;
    const int Q = ncols >> 2;                          // float4 column quads (ncols is a multiple of 64)
    const int QC = Q < EPI_THREADS ? Q : EPI_THREADS;  // quads handled concurrently; Q and EPI_THREADS are powers of two or Q > 256
    const int RL = EPI_THREADS / QC;                   // row lanes per quad (1 when Q >= 256)
    double* part = reinterpret_cast<double*>(scratch);                 // [RL][QC][8]
    double* tot = part + EPI_THREADS * 8;                              // [2][ncols]
    const int q0 = et % QC, rl = et / QC;
    for (int q = q0; q < Q; q += QC) {
      double acc[8] = {};
      if (rl < RL) {
        for (int r0 = rl; r0 < rows; r0 += RL * 4) {
          float4 v[4][2];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = r0 + u * RL;
            if (r < rows) {
              v[u][0] = __ldcg(reinterpret_cast<const float4*>(ep.stats + ((size_t)r * 2 + 0) * ncols) + q);
              v[u][1] = __ldcg(reinterpret_cast<const float4*>(ep.stats + ((size_t)r * 2 + 1) * ncols) + q);
            } else {
              v[u][0] = v[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            acc[0] += (double)v[u][0].x; acc[1] += (double)v[u][0].y; acc[2] += (double)v[u][0].z; acc[3] += (double)v[u][0].w;
            acc[4] += (double)v[u][1].x; acc[5] += (double)v[u][1].y; acc[6] += (double)v[u][1].z; acc[7] += (double)v[u][1].w;
          }
        }
      }
      if (RL > 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) part[((size_t)rl * QC + q0) * 8 + i] = acc[i];
        named_bar_sync(1, EPI_THREADS);                // RL > 1 implies Q <= 128: one pass of the q loop, the barrier is uniform
        if (rl == 0) {
          for (int l = 1; l < RL; ++l)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += part[((size_t)l * QC + q0) * 8 + i];
        }
      }
      if (rl == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          tot[q * 4 + i] = acc[i];
          tot[ncols + q * 4 + i] = acc[4 + i];
        }
      }
    }
    named_bar_sync(1, EPI_THREADS);
    const int C = ncols / ep.fin_groups;
    for (int c = et; c < C; c += EPI_THREADS) {
      double sum = 0.0, sq = 0.0;
      for (int g = 0; g < ep.fin_groups; ++g) {
        sum += tot[g * C + c];
        sq += tot[ncols + g * C + c];
      }
      const double mu = sum / ep.fin_count;
      double var = sq / ep.fin_count - mu * mu;
      if (var < 0.0) var = 0.0;
      ep.fin_mean[c] = (float)mu;
      ep.fin_rstd[c] = (float)(1.0 / sqrt(var + (double)ep.fin_eps));
      if (ep.fin_moving_mean) {
        const double unb = ep.fin_count > 1.0 ? var * (ep.fin_count / (ep.fin_count - 1.0)) : var;
        ep.fin_moving_mean[c] = ep.fin_momentum * ep.fin_moving_mean[c] + (1.f - ep.fin_momentum) * (float)mu;
        ep.fin_moving_var[c] = ep.fin_momentum * ep.fin_moving_var[c] + (1.f - ep.fin_momentum) * (float)unb;
      }
    }
    if (et == 0) *ep.fin_counter = 0u;                 // ready for the next launch
  }
};
