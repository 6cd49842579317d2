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
// CTA tile = MT horizontally adjacent 128-pixel tiles (MT = 2: a 16 x 16 super-tile, one {64, 18, 18} patch, two
// accumulators) x BLOCK_N columns.  The activation operand is nearly free (one patch feeds 9 taps), so what the L2 has to
// deliver per MMA clock is the WEIGHT tile: 8192 / (128 MT) bytes per clock per SM.  The first version (MT = 1, BLOCK_N =
// 256) asked for 64 B/clk/SM and ran into the ~8.5 kB/clk chip-wide L2->SM plateau (profiles/r01_summary.md); MT = 2
// with BLOCK_N = 128 halves that for the same 512 TMEM columns.
//
// Weights stream through a ring of per-(tap, channel-block) tiles; when the whole [BLOCK_N x 9 Cin] slice fits the
// ring (the 64-channel level-1 layers) it is loaded once per CTA and stays resident.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2-9 = epilogue (epilogue.cuh: bias, ReLU,
// optional folded inference BatchNorm, bf16 store through TMA, per-channel sum / sum-of-squares partials for training BN).
#include <stdlib.h>

#include "common.cuh"
#include "epilogue.cuh"

namespace {

constexpr int TW = 8, TH = 16;           // one accumulator = 16 (h) x 8 (w) pixels
constexpr int PH = TH + 2;

template <int MT>
struct Patch {
  static constexpr int TWS = TW * MT;                          // CTA tile width in pixels
  static constexpr int PW = TWS + 2;
  static constexpr int BYTES = PW * PH * 128;                  // 23040 (MT = 1) / 41472 (MT = 2)
  static constexpr int STRIDE = (BYTES + 1023) / 1024 * 1024;  // ring pitch (1024-aligned)
};

struct Conv3Params {
  CUtensorMap a_map[2];
  CUtensorMap b_map;
  CUtensorMap o_map[2];
  CUtensorMap red_map;       // RED: saved activation `a` of the BatchNorm'd tensor this dgrad writes the gradient of
  int nsrc;
  int cblk[2];
  int cblk_total;
  int H, W;
  int tiles_w, tiles_h;
  int n_tiles, total_tiles;
  int blocks_per_omap;
  EpiParams ep;
  int ncols;
  int b_resident;
  const void* w_base;        // host-side only: weight matrix [ncols][9 * Cin] for the tensor map
  const BnFin* fin;          // host-side only: BatchNorm to finalise in the last CTA (forward with statistics), or null
  const void* o_base[2];     // host-side only: output tensors behind o_map (the row-streaming kernel rebuilds the maps with its own box)
  int o_ch[2];
  const void* red_base;      // host-side only: tensor behind red_map
  int rows_total, strips_w;  // conv3_rows.cuh: row units (image, 128-pixel strip, row) of the launch; strips per image row
};

template <int BLOCK_N, int MT, int A_STAGES, int B_SLOTS, int OUT_BUFS, int RED, int CASEB = 0>
struct C3Smem {
  using E = EpiSmem<BLOCK_N, OUT_BUFS, RED, CASEB>;
  using P = Patch<MT>;
  static constexpr int B_BYTES = BLOCK_N * 128;
  static constexpr int OFF_B = A_STAGES * P::STRIDE;
  static constexpr int OFF_EPI = OFF_B + B_SLOTS * B_BYTES;
  static constexpr int OFF_BAR = OFF_EPI + E::TOTAL;
  static constexpr int NBAR = 2 * A_STAGES + 2 * B_SLOTS + 4;
  static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
  static constexpr int TOTAL = OFF_TMEM + 16 + 1024;
  static constexpr int STAGE_COLS = MT * BLOCK_N;              // TMEM columns per accumulator stage
  static constexpr int TMEM_COLS = 2 * STAGE_COLS;
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns");
};

template <int BLOCK_N, int MT, int A_STAGES, int B_SLOTS, int OUT_BUFS, int RED, int CASEB = 0>
__global__ void __launch_bounds__(64 + EPI_THREADS, 1) conv3_kernel(const __grid_constant__ Conv3Params p) {
  using L = C3Smem<BLOCK_N, MT, A_STAGES, B_SLOTS, OUT_BUFS, RED, CASEB>;
  using PT = Patch<MT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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
    if (RED) tma_prefetch_desc(&p.red_map);
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
      mbar_init(&tempty[a], EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, L::TMEM_COLS);
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
      const int w0 = (rem % p.tiles_w) * PT::TWS;
      int cbg = 0;
      for (int src = 0; src < p.nsrc; ++src) {
        for (int cb = 0; cb < p.cblk[src]; ++cb, ++cbg) {
          mbar_wait(&aempty[a_stage], a_phase ^ 1);
          mbar_expect_tx_e(&afull[a_stage], PT::BYTES);
          tma_load_4d_e(smem + a_stage * PT::STRIDE, &p.a_map[src], &afull[a_stage], cb * 64, w0 - 1, h0 - 1, img);
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
    // Kept lean on purpose: with 64-column tiles an MMA retires every 48 clocks and this warp's own instruction stream
    // was the bottleneck (profiles/r01g): descriptors are base + compile-time offsets, one elect per four MMAs, and
    // resident weights are waited for once (first tile) instead of once per tap.
    constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 0, 0);
    const uint32_t tmem_u = warp_uniform(tmem_base);
    const uint32_t smem_base_u = warp_uniform(smem_u32(smem));
    const uint64_t bdesc0 = make_smem_desc(smem_base_u + L::OFF_B, 16, 1024);
    int a_stage = 0, b_slot = 0;
    uint32_t a_phase = 0, b_phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    bool first = true;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_u + as * L::STAGE_COLS;
      for (int cbg = 0; cbg < p.cblk_total; ++cbg) {
        mbar_wait(&afull[a_stage], a_phase);
        if (resident && first)
          for (int tap = 0; tap < 9; ++tap) mbar_wait(&bfull[cbg * 9 + tap], 0u);
        tc_fence_after();
        const uint64_t adesc0 = make_smem_desc(smem_base_u + a_stage * PT::STRIDE, 16, PT::PW * 128);
        const uint64_t bdesc_cb = bdesc0 + (uint64_t)((cbg * 9 * L::B_BYTES) >> 4);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int dh = tap / 3, dw = tap % 3;           // already offset by +1 (patch origin is pixel (-1, -1))
          uint64_t bdesc;
          if (resident) {
            bdesc = bdesc_cb + (uint64_t)((tap * L::B_BYTES) >> 4);
          } else {
            mbar_wait(&bfull[b_slot], b_phase);
            tc_fence_after();
            bdesc = bdesc0 + (uint64_t)((b_slot * L::B_BYTES) >> 4);
          }
#pragma unroll
          for (int j = 0; j < MT; ++j)                    // the MT 16x8 pixel tiles of the super-tile share this weight tile
            tc_mma4_bf16_e(d_tmem + j * BLOCK_N, adesc0 + (uint64_t)(((dh * PT::PW + dw + j * TW) * 128) >> 4), bdesc, idesc,
                           (cbg | tap) != 0);
          if (!resident) {
            tc_commit_e(&bempty[b_slot]);
            if (++b_slot == B_SLOTS) { b_slot = 0; b_phase ^= 1; }
          }
        }
        tc_commit_e(&aempty[a_stage]);
        if (++a_stage == A_STAGES) { a_stage = 0; a_phase ^= 1; }
      }
      tc_commit_e(&tfull[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1;
      first = false;
    }
    __syncwarp();
  } else {
    // ================= epilogue (8 warps, epilogue.cuh) =================
    Epilogue<BLOCK_N, OUT_BUFS, TW, RED, CASEB> epi(smem + L::OFF_EPI, p.ep, tmem_base, tfull, tempty, threadIdx.x - 64, warp);
    epi.red_map = &p.red_map;
    epi.load_vectors(n_tile);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_tiles;
      const int img = m_tile / tiles_per_img;
      const int rem = m_tile - img * tiles_per_img;
      const int h0 = (rem / p.tiles_w) * TH;
      const int w0 = (rem % p.tiles_w) * PT::TWS;
#pragma unroll
      for (int j = 0; j < MT; ++j) {
        const int wj = w0 + j * TW;
        // RED == 2: the part after this one (its `a` tile is prefetched now)
        bool has_next = false;
        int nh0 = h0, nw0 = wj + TW, nimg = img;
        if (RED == 2) {
          if (j + 1 < MT) {
            has_next = true;
          } else if (tile + (int)gridDim.x < p.total_tiles) {
            const int nm = (tile + (int)gridDim.x) / p.n_tiles;
            nimg = nm / tiles_per_img;
            const int nrem = nm - nimg * tiles_per_img;
            nh0 = (nrem / p.tiles_w) * TH;
            nw0 = (nrem % p.tiles_w) * PT::TWS;
            has_next = true;
          }
        }
        epi.tile(h0, wj, [&](const uint8_t* blk, int b) {
          const int jb = n_tile * (BLOCK_N / 64) + b;
          const int map = jb / p.blocks_per_omap;
          tma_store_4d(&p.o_map[map], blk, (jb - map * p.blocks_per_omap) * 64, wj, h0, img);
        }, j * BLOCK_N, j == 0, j == MT - 1, L::STAGE_COLS, img, has_next, nh0, nw0, nimg);
      }
    }
    epi.finish(n_tile, blockIdx.x / p.n_tiles);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, L::TMEM_COLS);
  }
}

// conv3x3 forward / dgrad on CTA PAIRS (tcgen05 cta_group::2): the halo-patch implicit GEMM of igemm_conv3.cu with M = 256 per MMA.
// Two CTAs of a cluster (two SMs of one TPC) take two horizontally independent pixel super-tiles and the SAME 128 output columns;
// each loads its own activation patch and HALF of every weight tile, the leader issues tcgen05.mma.cta_group::2 for both, and each CTA
// drains its own 128 TMEM lanes through the unchanged epilogue.  Per SM and per 64-channel block the shared-memory port then serves
// 72 x 6 KB of operand reads + 115 KB of TMA writes instead of 72 x 8 KB + 188 KB: ~119 B/clk against the 128 B/clk port, where the
// single-CTA kernel asks for ~166 B/clk and is port-bound.  Selected for the 128-column tiles by UB_CONV3_2CTA=1.
//
// Protocol (every barrier exists at the same shared-memory offset in both CTAs):
//   afull / bfull   used in the LEADER only: its producer arms them with the bytes of BOTH CTAs; the peer's TMA loads complete_tx on
//                   the leader's barrier (cp.async.bulk.tensor ... cta_group::2 with the peer bit of the barrier address cleared)
//   aempty / bempty one arrive in EACH CTA from the leader's tcgen05.commit ... multicast::cluster (mask 0b11): each producer refills
//                   its own shared memory
//   tfull           multicast commit as well: each CTA's epilogue waits on its own copy
//   tempty          LEADER only, 2 x EPI_WARPS arrivals: the peer's epilogue warps arrive remotely (mapa + mbarrier.arrive ... cluster)
template <int BLOCK_N, int MT, int A_STAGES, int B_SLOTS, int OUT_BUFS, int RED, int CASEB = 0>
struct C3PSmem {
  using E = EpiSmem<BLOCK_N, OUT_BUFS, RED, CASEB>;
  using P = Patch<MT>;
  static constexpr int B_BYTES = (BLOCK_N / 2) * 128;      // this CTA's half of a weight tile
  static constexpr int OFF_B = A_STAGES * P::STRIDE;
  static constexpr int OFF_EPI = OFF_B + B_SLOTS * B_BYTES;
  static constexpr int OFF_BAR = OFF_EPI + E::TOTAL;
  static constexpr int NBAR = 2 * A_STAGES + 2 * B_SLOTS + 4;
  static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
  static constexpr int TOTAL = OFF_TMEM + 16 + 1024;
  static constexpr int STAGE_COLS = MT * BLOCK_N;              // TMEM columns per accumulator stage
  static constexpr int TMEM_COLS = 2 * STAGE_COLS;
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns");
};

template <int BLOCK_N, int MT, int A_STAGES, int B_SLOTS, int OUT_BUFS, int RED, int CASEB = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + EPI_THREADS, 1) conv3_pair_kernel(const __grid_constant__ Conv3Params p) {
  using L = C3PSmem<BLOCK_N, MT, A_STAGES, B_SLOTS, OUT_BUFS, RED, CASEB>;
  using PT = Patch<MT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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
    if (RED) tma_prefetch_desc(&p.red_map);
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
      mbar_init(&tempty[a], 2 * EPI_WARPS);        // epilogue warps of both CTAs (used in the leader)
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, L::TMEM_COLS);      // warp 1 of BOTH CTAs, same shared-memory slot
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int rank = (int)cluster_ctarank();          // 0 = leader (issues the MMAs of the pair)
  const int cid = blockIdx.x >> 1, NC = gridDim.x >> 1;
  const int n_tile = cid % p.n_tiles;               // host guarantees NC % n_tiles == 0: fixed per cluster
  const bool leader = rank == 0;
  const bool resident = p.b_resident != 0;

  if (warp == 0) {
    // ================= TMA producer (all lanes run the loop, one elected lane issues) =================
    int a_stage = 0, b_slot = 0;
    uint32_t a_phase = 0, b_phase = 0;
    bool first = true;
    for (int tile = cid; tile < p.total_tiles; tile += NC) {
      const int m_tile = 2 * (tile / p.n_tiles) + rank;
      const int img = m_tile / tiles_per_img;
      const int rem = m_tile - img * tiles_per_img;
      const int h0 = (rem / p.tiles_w) * TH;
      const int w0 = (rem % p.tiles_w) * PT::TWS;
      int cbg = 0;
      for (int src = 0; src < p.nsrc; ++src) {
        for (int cb = 0; cb < p.cblk[src]; ++cb, ++cbg) {
          mbar_wait(&aempty[a_stage], a_phase ^ 1);
          if (leader) mbar_expect_tx_e(&afull[a_stage], 2 * PT::BYTES);          // this CTA's patch and the peer's
          tma_load_4d_2sm_e(smem + a_stage * PT::STRIDE, &p.a_map[src], &afull[a_stage], cb * 64, w0 - 1, h0 - 1, img);
          if (++a_stage == A_STAGES) { a_stage = 0; a_phase ^= 1; }
          if (resident && !first) continue;
          for (int tap = 0; tap < 9; ++tap) {
            const int slot = resident ? cbg * 9 + tap : b_slot;
            if (!resident) mbar_wait(&bempty[slot], b_phase ^ 1);
            if (leader) mbar_expect_tx_e(&bfull[slot], 2 * L::B_BYTES);
            tma_load_2d_2sm_e(smem + L::OFF_B + slot * L::B_BYTES, &p.b_map, &bfull[slot], (tap * p.cblk_total + cbg) * 64,
                              n_tile * BLOCK_N + rank * (BLOCK_N / 2));
            if (!resident && ++b_slot == B_SLOTS) { b_slot = 0; b_phase ^= 1; }
          }
        }
      }
      first = false;
    }
    __syncwarp();
  } else if (warp == 1 && leader) {
    // ================= MMA issuer (all lanes run the loop, one elected lane issues) =================
    // Kept lean on purpose: with 64-column tiles an MMA retires every 48 clocks and this warp's own instruction stream
    // was the bottleneck (profiles/r01g): descriptors are base + compile-time offsets, one elect per four MMAs, and
    // resident weights are waited for once (first tile) instead of once per tap.
    constexpr uint32_t idesc = make_idesc_bf16(256, BLOCK_N, 0, 0);      // M = 256: 128 rows in each CTA
    const uint32_t tmem_u = warp_uniform(tmem_base);
    const uint32_t smem_base_u = warp_uniform(smem_u32(smem));
    const uint64_t bdesc0 = make_smem_desc(smem_base_u + L::OFF_B, 16, 1024);
    int a_stage = 0, b_slot = 0;
    uint32_t a_phase = 0, b_phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    bool first = true;
    for (int tile = cid; tile < p.total_tiles; tile += NC) {
      mbar_wait(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_u + as * L::STAGE_COLS;
      for (int cbg = 0; cbg < p.cblk_total; ++cbg) {
        mbar_wait(&afull[a_stage], a_phase);
        if (resident && first)
          for (int tap = 0; tap < 9; ++tap) mbar_wait(&bfull[cbg * 9 + tap], 0u);
        tc_fence_after();
        const uint64_t adesc0 = make_smem_desc(smem_base_u + a_stage * PT::STRIDE, 16, PT::PW * 128);
        const uint64_t bdesc_cb = bdesc0 + (uint64_t)((cbg * 9 * L::B_BYTES) >> 4);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int dh = tap / 3, dw = tap % 3;           // already offset by +1 (patch origin is pixel (-1, -1))
          uint64_t bdesc;
          if (resident) {
            bdesc = bdesc_cb + (uint64_t)((tap * L::B_BYTES) >> 4);
          } else {
            mbar_wait(&bfull[b_slot], b_phase);
            tc_fence_after();
            bdesc = bdesc0 + (uint64_t)((b_slot * L::B_BYTES) >> 4);
          }
#pragma unroll
          for (int j = 0; j < MT; ++j)                    // the MT 16x8 pixel tiles of the super-tile share this weight tile
            tc_mma4_bf16_2sm_e(d_tmem + j * BLOCK_N, adesc0 + (uint64_t)(((dh * PT::PW + dw + j * TW) * 128) >> 4), bdesc, idesc,
                           (cbg | tap) != 0);
          if (!resident) {
            tc_commit_2sm_e(&bempty[b_slot]);
            if (++b_slot == B_SLOTS) { b_slot = 0; b_phase ^= 1; }
          }
        }
        tc_commit_2sm_e(&aempty[a_stage]);
        if (++a_stage == A_STAGES) { a_stage = 0; a_phase ^= 1; }
      }
      tc_commit_2sm_e(&tfull[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1;
      first = false;
    }
    __syncwarp();
  } else if (warp >= 2) {
    // ================= epilogue (8 warps in each CTA, epilogue.cuh) =================
    Epilogue<BLOCK_N, OUT_BUFS, TW, RED, CASEB> epi(smem + L::OFF_EPI, p.ep, tmem_base, tfull, tempty, threadIdx.x - 64, warp);
    epi.red_map = &p.red_map;
    epi.pair_rank = rank;
    epi.load_vectors(n_tile);
    for (int tile = cid; tile < p.total_tiles; tile += NC) {
      const int m_tile = 2 * (tile / p.n_tiles) + rank;
      const int img = m_tile / tiles_per_img;
      const int rem = m_tile - img * tiles_per_img;
      const int h0 = (rem / p.tiles_w) * TH;
      const int w0 = (rem % p.tiles_w) * PT::TWS;
#pragma unroll
      for (int j = 0; j < MT; ++j) {
        const int wj = w0 + j * TW;
        // RED == 2: the part after this one (its `a` tile is prefetched now)
        bool has_next = false;
        int nh0 = h0, nw0 = wj + TW, nimg = img;
        if (RED == 2) {
          if (j + 1 < MT) {
            has_next = true;
          } else if (tile + NC < p.total_tiles) {
            const int nm = 2 * ((tile + NC) / p.n_tiles) + rank;
            nimg = nm / tiles_per_img;
            const int nrem = nm - nimg * tiles_per_img;
            nh0 = (nrem / p.tiles_w) * TH;
            nw0 = (nrem % p.tiles_w) * PT::TWS;
            has_next = true;
          }
        }
        epi.tile(h0, wj, [&](const uint8_t* blk, int b) {
          const int jb = n_tile * (BLOCK_N / 64) + b;
          const int map = jb / p.blocks_per_omap;
          tma_store_4d(&p.o_map[map], blk, (jb - map * p.blocks_per_omap) * 64, wj, h0, img);
        }, j * BLOCK_N, j == 0, j == MT - 1, L::STAGE_COLS, img, has_next, nh0, nw0, nimg);
      }
    }
    epi.finish(n_tile, 2 * (cid / p.n_tiles) + rank);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // neither CTA leaves (or frees TMEM) while the other may still signal its barriers or run pair MMAs
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, L::TMEM_COLS);
  }
}

template <int BLOCK_N, int MT, int A_STAGES, int B_SLOTS, int OUT_BUFS, int RED = 0, int CASEB = 0>
int launch_c3_pair(Conv3Params& p, const void* const* a_base, const int* a_ch, int n_img, cudaStream_t stream, bool* taken) {
  using L = C3PSmem<BLOCK_N, MT, A_STAGES, B_SLOTS, OUT_BUFS, RED, CASEB>;
  using PT = Patch<MT>;
  static_assert(L::TOTAL <= 232448, "smem budget");
  *taken = false;
  const int n_tiles = p.ncols / BLOCK_N;
  const int tiles_w = (p.W + PT::TWS - 1) / PT::TWS, tiles_h = (p.H + TH - 1) / TH;
  const long long m_tiles = (long long)n_img * tiles_w * tiles_h;
  const int pairs_avail = ub_num_sms() / 2;
  if (m_tiles % 2 != 0 || n_tiles > pairs_avail) return UB_OK;          // an odd pixel-tile count or too many column tiles: single-CTA kernel
  *taken = true;
  auto kern = conv3_pair_kernel<BLOCK_N, MT, A_STAGES, B_SLOTS, OUT_BUFS, RED, CASEB>;
  static bool attr_done = false;
  if (!attr_done) {
    UB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_done = true;
  }
  int rc;
  for (int i = 0; i < p.nsrc; ++i)
    if ((rc = ub_tmap_act4d(&p.a_map[i], a_base[i], a_ch[i], p.W, p.H, n_img, (long long)a_ch[i] * 2, (long long)p.W * a_ch[i] * 2,
                            (long long)p.H * p.W * a_ch[i] * 2, PT::PW, PH)))
      return rc;
  if ((rc = ub_tmap_mat2d(&p.b_map, p.w_base, p.ncols, 9ll * p.cblk_total * 64, BLOCK_N / 2))) return rc;      // each CTA loads half a tile
  p.n_tiles = n_tiles;
  p.tiles_w = tiles_w;
  p.tiles_h = tiles_h;
  const long long total = (m_tiles / 2) * n_tiles;                       // pair tiles: two pixel super-tiles x one column tile
  UB_CHECK_SHAPE(total > 0 && total < (1ll << 30), "conv3 (pair): tile count out of range");
  p.total_tiles = (int)total;
  p.b_resident = (9 * p.cblk_total <= B_SLOTS) ? 1 : 0;                 // 64-channel layers: the whole (half) weight slice stays in shared memory
  long long clusters = (long long)(pairs_avail / n_tiles) * n_tiles;
  if (clusters > total) clusters = total;                                // total is a multiple of n_tiles
  const long long grid = 2 * clusters;
  UB_CHECK_SHAPE(grid / n_tiles <= UB_STATS_ROWS, "conv3 (pair): stats rows");
  const BnFin* fin = p.fin;
  p.fin = nullptr;
  const bool fused = fin && p.ep.stats && epi_fin_bytes(p.ncols) <= OUT_BUFS * L::E::OUT_BYTES;
  if (fused) epi_set_fin(p.ep, *fin, (int)(grid / n_tiles));
  else if (p.ep.stats) UB_CUDA(cudaMemsetAsync(p.ep.stats, 0, sizeof(float) * UB_STATS_ROWS * 2 * p.ncols, stream));
  if (RED) UB_CUDA(cudaMemsetAsync(p.ep.red_out, 0, sizeof(float) * UB_STATS_ROWS * 2 * p.ep.red_ncols, stream));
  kern<<<(int)grid, 64 + EPI_THREADS, L::TOTAL, stream>>>(p);           // __cluster_dims__(2, 1, 1)
  UB_LAUNCH_CHECK();
  if (fin && !fused)
    return ub_bn_finalize(p.ep.stats, p.ncols, fin->groups, (long long)fin->count, fin->mean, fin->rstd, fin->moving_mean, fin->moving_var,
                          fin->momentum, fin->eps, stream);
  return UB_OK;
}

static bool use_pairs() {
  static int v = -1;
  if (v < 0) {
    // CTA-pair (cta_group::2) kernel: parity green and 5-10 % faster sustained on the 128-column tiles (profiles/r02_ab_runs.md);
    // UB_CONV3_2CTA=0 selects the single-CTA kernels for A/B runs
    const char* e = getenv("UB_CONV3_2CTA");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
static int use_pairs256() {          // 1 = forced on, 0 = forced off, -1 = by shape
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("UB_CONV3_PAIR256");
    v = !e ? -1 : (e[0] == '1' ? 1 : 0);
  }
  return v;
}
static bool use_pairs64() {
  static int v = -1;
  if (v < 0) {
    // ... also for the 64-output-channel layers: parity green, but SLOWER (enc1b forward 718 -> 660 TFLOP/s sustained, dgrad 805 -> 738:
    // an M = 256 x N = 64 pair MMA does not retire faster than two M = 128 x N = 64 ones, and the pair protocol adds latency): off
    const char* e = getenv("UB_CONV3_2CTA_64");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

template <int BLOCK_N, int MT, int A_STAGES, int B_SLOTS, int OUT_BUFS, int RED = 0, int CASEB = 0>
int launch_c3(Conv3Params& p, const void* const* a_base, const int* a_ch, int n_img, cudaStream_t stream) {
  using L = C3Smem<BLOCK_N, MT, A_STAGES, B_SLOTS, OUT_BUFS, RED, CASEB>;
  using PT = Patch<MT>;
  static_assert(L::TOTAL <= 232448, "smem budget");
  auto kern = conv3_kernel<BLOCK_N, MT, A_STAGES, B_SLOTS, OUT_BUFS, RED, CASEB>;
  static bool attr_done = false;
  if (!attr_done) {
    UB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_done = true;
  }
  int rc;
  for (int i = 0; i < p.nsrc; ++i)
    if ((rc = ub_tmap_act4d(&p.a_map[i], a_base[i], a_ch[i], p.W, p.H, n_img, (long long)a_ch[i] * 2, (long long)p.W * a_ch[i] * 2,
                            (long long)p.H * p.W * a_ch[i] * 2, PT::PW, PH)))
      return rc;
  if ((rc = ub_tmap_mat2d(&p.b_map, p.w_base, p.ncols, 9ll * p.cblk_total * 64, BLOCK_N))) return rc;
  p.n_tiles = p.ncols / BLOCK_N;
  p.tiles_w = (p.W + PT::TWS - 1) / PT::TWS;
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
  const BnFin* fin = p.fin;
  p.fin = nullptr;
  const bool fused = fin && p.ep.stats && epi_fin_bytes(p.ncols) <= OUT_BUFS * L::E::OUT_BYTES;
  if (fused) epi_set_fin(p.ep, *fin, (int)(grid / p.n_tiles));           // the last CTA reads exactly the rows this grid writes: no zero-fill
  else if (p.ep.stats) UB_CUDA(cudaMemsetAsync(p.ep.stats, 0, sizeof(float) * UB_STATS_ROWS * 2 * p.ncols, stream));
  if (RED) UB_CUDA(cudaMemsetAsync(p.ep.red_out, 0, sizeof(float) * UB_STATS_ROWS * 2 * p.ep.red_ncols, stream));
  kern<<<(int)grid, 64 + EPI_THREADS, L::TOTAL, stream>>>(p);
  UB_LAUNCH_CHECK();
  if (fin && !fused)
    return ub_bn_finalize(p.ep.stats, p.ncols, fin->groups, (long long)fin->count, fin->mean, fin->rstd, fin->moving_mean, fin->moving_var,
                          fin->momentum, fin->eps, stream);
  return UB_OK;
}

#include "conv3_rows.cuh"

int launch(Conv3Params& p, const void* const* a_base, const int* a_ch, int n_img, cudaStream_t stream, bool bias_cases = false) {
  p.cblk_total = 0;
  for (int i = 0; i < p.nsrc; ++i) p.cblk_total += p.cblk[i];
  UB_CHECK_SHAPE(p.ncols % 64 == 0 && p.cblk_total > 0, "conv3: columns must be a multiple of 64");
  p.ep.ncols = p.ncols;
  p.ep.bias_mod = p.ncols;
  if (use_rows() >= 3 && p.ncols == 128 && p.cblk_total == 1 && p.o_ch[0] == 64 && p.o_ch[1] == 64 && !bias_cases && !p.ep.stats && !p.ep.bias &&
      !p.ep.post_scale && (!p.ep.red_out || (p.ep.red_blk_begin == 1 && p.ep.red_ncols == 64)) && rows_width_ok(p.W)) {
    // dgrad of a 64-channel convolution over a concat of 64 + 64 channels (dec1a): two row-streaming launches, one per output tensor --
    // each reads dz once more than a 128-column tile would, and runs ~10 % faster than the 128-column pair kernel whose K = 576 loop is
    // too short for its operand traffic; the second launch carries the fused BatchNorm-backward reduction (dx1 = dL/dy of up1)
    Conv3Params q = p;
    q.ncols = q.ep.ncols = q.ep.bias_mod = 64;
    q.ep.red_out = nullptr;
    q.ep.red_mean = q.ep.red_rstd = nullptr;
    q.ep.red_ncols = q.ep.red_blk_begin = 0;
    int rc = launch_c3_rows<1, 6, 2, 0>(q, a_base, a_ch, n_img, stream);
    if (rc) return rc;
    Conv3Params r = p;
    r.ncols = r.ep.ncols = r.ep.bias_mod = 64;
    r.w_base = static_cast<const char*>(p.w_base) + (size_t)64 * 9 * 64 * 2;      // rows 64 .. 127 of the [128][9 * 64] dgrad weight matrix
    r.o_base[0] = p.o_base[1];
    r.o_ch[0] = p.o_ch[1];
    r.ep.red_blk_begin = 0;
    return p.ep.red_out ? launch_c3_rows<1, 4, 2, 0, 2>(r, a_base, a_ch, n_img, stream) : launch_c3_rows<1, 6, 2, 0>(r, a_base, a_ch, n_img, stream);
  }
  if (use_rows() >= 1 && p.ncols == 64 && p.ep.red_out && p.o_ch[0] == 64 && p.ep.red_ncols == 64 && p.cblk_total == 1 && rows_width_ok(p.W))
    return launch_c3_rows<1, 4, 2, 0, 2>(p, a_base, a_ch, n_img, stream);      // 64 -> 64 dgrad with the fused BatchNorm-backward reduction
  if (use_rows() >= 1 && p.ncols == 64 && !p.ep.red_out && p.o_ch[0] == 64 && rows_width_ok(p.W)) {
    // 64 output channels and rows that fill their 128-pixel strips (level 1 at the benchmark shapes, the 1216-pixel inference tiles): the
    // row-streaming kernel (conv3_rows.cuh)
    if (p.cblk_total == 1)
      return bias_cases ? launch_c3_rows<1, 6, 2, 1>(p, a_base, a_ch, n_img, stream) : launch_c3_rows<1, 6, 2, 0>(p, a_base, a_ch, n_img, stream);
    if (p.cblk_total == 2 && use_rows() >= 2)
      return bias_cases ? launch_c3_rows<2, 3, 1, 1>(p, a_base, a_ch, n_img, stream) : launch_c3_rows<2, 3, 1, 0>(p, a_base, a_ch, n_img, stream);
  }
  if (use_pairs() && !p.ep.red_out && p.ncols % 256 == 0 && (use_pairs256() == 1 || (use_pairs256() < 0 && p.ncols == 512))) {
    // 128 pixels x 256 columns per CTA, N = 256 per pair MMA (each CTA holds 128 columns of the weight tile): 8 KB of operand reads per
    // 128-clock MMA instead of 6 KB per 64-clock one, at the price of a weight tile per 128 instead of 256 pixels.  Measured sustained
    // (profiles/r02_ab_runs.md, block I): 512-column layers +6 % (enc4b forward 1262 -> 1347, dgrad 1300 -> 1375 TFLOP/s), 256- and
    // 1024-column layers +-1 %: default for 512 columns; UB_CONV3_PAIR256=1 / 0 forces it on for every multiple of 256 / off.
    bool taken = false;
    int rc = bias_cases ? launch_c3_pair<256, 1, 2, 4, 1, 0, 1>(p, a_base, a_ch, n_img, stream, &taken)
                        : launch_c3_pair<256, 1, 2, 4, 1>(p, a_base, a_ch, n_img, stream, &taken);
    if (taken || rc) return rc;
  }
  if (use_pairs() && p.ncols % 128 == 0) {
    bool taken = false;
    int rc;
    if (bias_cases) rc = launch_c3_pair<128, 2, 2, 8, 2, 0, 1>(p, a_base, a_ch, n_img, stream, &taken);
    else if (p.ep.red_out) rc = launch_c3_pair<128, 2, 2, 8, 1, 1>(p, a_base, a_ch, n_img, stream, &taken);
    else rc = launch_c3_pair<128, 2, 2, 8, 2>(p, a_base, a_ch, n_img, stream, &taken);
    if (taken || rc) return rc;
  } else if (use_pairs() && use_pairs64() && !p.ep.red_out && p.cblk_total <= 2) {
    // 64 output channels (level 1): M = 256 x N = 64 per MMA, each CTA holds 32 columns of the resident weights (5 KB instead of 6 KB of
    // operand reads per MMA: the layer is bound by the shared-memory port)
    bool taken = false;
    int rc;
    if (bias_cases) rc = p.cblk_total == 1 ? launch_c3_pair<64, 2, 2, 9, 2, 0, 1>(p, a_base, a_ch, n_img, stream, &taken)
                                           : launch_c3_pair<64, 1, 2, 18, 2, 0, 1>(p, a_base, a_ch, n_img, stream, &taken);
    else rc = p.cblk_total == 1 ? launch_c3_pair<64, 2, 2, 9, 2>(p, a_base, a_ch, n_img, stream, &taken)
                                : launch_c3_pair<64, 1, 2, 18, 2>(p, a_base, a_ch, n_img, stream, &taken);
    if (taken || rc) return rc;
  }
  if (bias_cases) {            // forward of a BatchNorm-folded input: 9-case border bias (H, W >= 2 checked by the caller)
    if (p.ncols % 128 == 0) return launch_c3<128, 2, 2, 4, 2, 0, 1>(p, a_base, a_ch, n_img, stream);
    if (p.cblk_total == 1) return launch_c3<64, 2, 2, 9, 2, 0, 1>(p, a_base, a_ch, n_img, stream);
    return launch_c3<64, 1, 2, 18, 2, 0, 1>(p, a_base, a_ch, n_img, stream);
  }
  static int v1 = -1;
  if (v1 < 0) {
    const char* e = getenv("UB_CONV3_V1");       // A/B switch for tools/bench_layers.py: the round-1 128-pixel tiles
    v1 = (e && e[0] == '1') ? 1 : 0;
  }
  if (v1) {
    if (p.ncols % 256 == 0) return launch_c3<256, 1, 2, 3, 1>(p, a_base, a_ch, n_img, stream);
    if (p.ncols % 128 == 0) return launch_c3<128, 1, 2, 6, 2>(p, a_base, a_ch, n_img, stream);
    return launch_c3<64, 1, 2, 18, 2>(p, a_base, a_ch, n_img, stream);
  }
  if (p.ep.red_out) {          // dgrad fused with the backward-BatchNorm reduction of the tensor it differentiates
    if (p.ncols % 128 == 0) return launch_c3<128, 2, 2, 4, 1, 1>(p, a_base, a_ch, n_img, stream);
    if (p.cblk_total == 1) return launch_c3<64, 2, 2, 9, 2, 2>(p, a_base, a_ch, n_img, stream);       // room for two `a` buffers: prefetch
    return launch_c3<64, 1, 2, 18, 1, 1>(p, a_base, a_ch, n_img, stream);
  }
  if (p.ncols % 128 == 0) return launch_c3<128, 2, 2, 4, 2>(p, a_base, a_ch, n_img, stream);       // 256 pixels x 128 columns
  static int v64 = -1;
  if (v64 < 0) {
    const char* e = getenv("UB_CONV3_64");        // A/B switch for tools/bench_layers.py: pipeline shape of the 64 -> 64 layers
    v64 = e ? atoi(e) : 0;
  }
  if (p.cblk_total == 1 && v64 == 1) return launch_c3<64, 2, 3, 9, 1>(p, a_base, a_ch, n_img, stream);   // 3 patches in flight, one staging buffer
  if (p.cblk_total == 1 && v64 == 2) return launch_c3<64, 1, 4, 9, 2>(p, a_base, a_ch, n_img, stream);   // 128-pixel tiles, 4 patches in flight
  if (p.cblk_total == 1 && v64 == 3) return launch_c3<64, 1, 5, 9, 2>(p, a_base, a_ch, n_img, stream);   // 128-pixel tiles, 5 patches in flight
  if (p.cblk_total == 1 && v64 == 4) return launch_c3<64, 2, 3, 4, 2>(p, a_base, a_ch, n_img, stream);   // streamed weights, 3 patches in flight
  if (p.cblk_total == 1 && v64 == 5) return launch_c3<64, 2, 4, 3, 1>(p, a_base, a_ch, n_img, stream);   // streamed weights, 4 patches in flight
  static int v64b = -1;
  if (v64b < 0) {
    const char* e = getenv("UB_CONV3_64B");       // same for the 128 -> 64 layer (dec1a)
    v64b = e ? atoi(e) : 0;
  }
  if (p.cblk_total == 2 && v64b == 1) return launch_c3<64, 2, 3, 6, 2>(p, a_base, a_ch, n_img, stream);  // 256-pixel super-tiles, streamed weights
  if (p.cblk_total == 2 && v64b == 2) return launch_c3<64, 2, 2, 9, 2>(p, a_base, a_ch, n_img, stream);
  if (p.cblk_total == 1) return launch_c3<64, 2, 2, 9, 2>(p, a_base, a_ch, n_img, stream);         // 64 -> 64: weights resident, 256-pixel super-tiles
  return launch_c3<64, 1, 2, 18, 2>(p, a_base, a_ch, n_img, stream);                               // 128 -> 64 (dec1a): weights resident
}

int out_map(CUtensorMap* m, const void* base, int C, int W, int H, int N) {
  return ub_tmap_act4d(m, base, C, W, H, N, (long long)C * 2, (long long)W * C * 2, (long long)H * W * C * 2, TW, TH);
}

}  // namespace

// Called by the extern "C" entry points in igemm_fwd.cu.
int ub_conv3_halo_fwd(const void* x0, int C0, const void* x1, int C1, const void* w, const float* bias, const float* post_scale,
                      const float* post_shift, void* out, float* stats, int N, int H, int W, int Cout, int relu, cudaStream_t stream,
                      int bias_cases, const BnFin* fin) {
  Conv3Params p;
  memset(&p, 0, sizeof(p));
  p.fin = fin;
  int rc;
  const void* a_base[2] = {x0, x1};
  const int a_ch[2] = {C0, C1};
  p.nsrc = C1 > 0 ? 2 : 1;
  p.cblk[0] = C0 / 64;
  p.cblk[1] = C1 / 64;
  p.w_base = w;
  if ((rc = out_map(&p.o_map[0], out, Cout, W, H, N))) return rc;
  p.o_base[0] = out;
  p.o_ch[0] = Cout;
  p.H = p.ep.H = H;
  p.W = p.ep.W = W;
  p.blocks_per_omap = Cout / 64;
  p.ep.bias = bias;
  p.ep.post_scale = post_scale;
  p.ep.post_shift = post_shift;
  p.ep.relu = relu;
  p.ep.stats = stats;
  p.ncols = Cout;
  if (bias_cases) UB_CHECK_SHAPE(H >= 2 && W >= 2, "conv3 with a border-case bias needs H, W >= 2 (got %d x %d)", H, W);
  return launch(p, a_base, a_ch, N, stream, bias_cases != 0);
}

int ub_conv3_halo_dgrad(const void* dz, int Cout, const void* w_t, void* dx0, int C0, void* dx1, int C1, int N, int H, int W,
                        const void* red_a, const float* red_mean, const float* red_rstd, float* red_partial, cudaStream_t stream) {
  Conv3Params p;
  memset(&p, 0, sizeof(p));
  int rc;
  const void* a_base[2] = {dz, nullptr};
  const int a_ch[2] = {Cout, 0};
  p.nsrc = 1;
  p.cblk[0] = Cout / 64;
  p.w_base = w_t;
  if ((rc = out_map(&p.o_map[0], dx0, C0, W, H, N))) return rc;
  if (C1 > 0 && (rc = out_map(&p.o_map[1], dx1, C1, W, H, N))) return rc;
  p.o_base[0] = dx0;
  p.o_ch[0] = C0;
  p.o_base[1] = dx1;
  p.o_ch[1] = C1;
  p.H = p.ep.H = H;
  p.W = p.ep.W = W;
  p.blocks_per_omap = C0 / 64;
  p.ncols = C0 + C1;
  if (red_partial) {
    // the BatchNorm'd tensor is the LAST output (dx1 of a concat dgrad, else dx0)
    const int cred = C1 > 0 ? C1 : C0;
    if ((rc = out_map(&p.red_map, red_a, cred, W, H, N))) return rc;
    p.red_base = red_a;
    p.ep.red_mean = red_mean;
    p.ep.red_rstd = red_rstd;
    p.ep.red_out = red_partial;
    p.ep.red_blk_begin = C1 > 0 ? C0 / 64 : 0;
    p.ep.red_ncols = cred;
  }
  return launch(p, a_base, a_ch, N, stream);
}
