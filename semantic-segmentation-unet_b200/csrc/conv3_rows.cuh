// conv3x3 ('same') forward / dgrad for the 64-OUTPUT-channel layers (level 1 of the U-Net: enc1b, dec1a, dec1b; UNet/model.py:88-90,
// :130-134) as a ROW-STREAMING implicit GEMM on tcgen05/TMEM (sm_100a).  Included by igemm_conv3.cu inside its anonymous namespace.
//
// Why a second kernel: with 64 output columns the halo-patch kernel issues 128 x 64 x 16 MMAs that read 6 KB of operands (A 4 KB +
// B 2 KB) through the 128 B/clk shared-memory port for 32 clocks of tensor work -- port-bound at 48 clocks per MMA ("What
// bounds the level-1 convolutions").  Here ONE A view feeds the three filter ROWS of a column offset at once:
//
//   M = 128 consecutive pixels of input row y (shifted by dw - 1), K = 64 input channels,
//   N = 192 = [W(dh=2, dw) | W(dh=1, dw) | W(dh=0, dw)] x 64 output channels,
//
// and the three 64-column groups of D are the accumulators of the output rows y - 1, y, y + 1, which sit side by side in a RING of
// eight 64-column TMEM slots (output row r lives in slot r mod 8).  An output row is complete after the MMAs of input row r + 1; the
// epilogue drains it, stores zeros back (every MMA accumulates, so there is no first-touch case inside an N = 192 window) and hands the
// slot to the row eight further down.  Per 128 x 192 x 16 MMA the port serves 10 KB in 96 tensor clocks (0.83) instead of 6 KB in 32
// (1.5); activation rows enter shared memory exactly once ({64 ch, 130 px} TMA boxes, no vertical halo re-read).
//
// Work = "row units" (image, 128-pixel column strip, row), split evenly over the CTAs in that order; a CTA walks its range as
// SEGMENTS of consecutive rows of one strip and loads one halo row above and below each segment.  The last strip of a row may be
// ragged (TMA loads zero-fill, TMA stores clip, the statistics mask): rows_width_ok().
// Warp roles as in igemm_conv3.cu: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM alloc), warps 2-9 epilogue (epilogue.cuh: bias / 9-case
// border bias, ReLU, optional inference affine, bf16 staging + TMA store, forward statistics with the last-CTA BatchNorm finalise).
// RED = 2 (64 -> 64 dgrads): the epilogue also accumulates the BatchNorm-backward sums of the tensor it writes from the matching row of
// the saved activation, TMA-loaded one row ahead (ub_conv3x3_dgrad_bnred).
// Barriers: afull / aempty per activation-row stage, wfull (weights, once), rfull[8] (tcgen05.commit: output row complete),
// rempty[8] (8 warp arrivals: slot drained and zeroed; armed once at start after the initial zero fill of all 512 columns).
// The index arithmetic is replayed in numpy against a direct convolution in tests/test_rows_schedule_cpu.py.
#pragma once

constexpr int RW = 128;                      // output pixels per row unit = MMA M
constexpr int RPW = RW + 2;                  // halo'd input row
constexpr int ROW_BYTES = RPW * 128;         // 16640
constexpr int ROW_STRIDE = 17 * 1024;        // ring pitch (1024-aligned)
constexpr int RING = 8;                      // TMEM slots of 64 columns
constexpr int WT_BYTES = 64 * 128;           // one (tap, channel block) weight tile: 64 output channels x 64 input channels

template <int CBLK, int A_STAGES, int OUT_BUFS, int CASEB, int RED = 0>
struct C3RSmem {
  using E = EpiSmem<64, OUT_BUFS, RED, CASEB>;
  static constexpr int OFF_W = 0;                                   // [(cbg, dw)][group g = 2 - dh][64 co][64 ci], resident
  static constexpr int OFF_A = 9 * CBLK * WT_BYTES;
  static constexpr int OFF_EPI = OFF_A + A_STAGES * ROW_STRIDE;
  static constexpr int OFF_BAR = OFF_EPI + E::TOTAL;
  static constexpr int NBAR = 2 * A_STAGES + 1 + 2 * RING;
  static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
  static constexpr int TOTAL = OFF_TMEM + 16 + 1024;
  static_assert(OFF_EPI % 1024 == 0 && OFF_BAR % 8 == 0, "alignment");
};

__device__ __forceinline__ void tmem_st_zero_32x32(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// this CTA's range of row units and the walk over its segments (identical in the three warp roles)
struct RowWalk {
  long long u, u1;
  int H, strips_w;
  __device__ __forceinline__ RowWalk(const Conv3Params& p) : H(p.H), strips_w(p.strips_w) {
    u = (long long)p.rows_total * blockIdx.x / gridDim.x;
    u1 = (long long)p.rows_total * (blockIdx.x + 1) / gridDim.x;
  }
  // next segment: rows [hb, hb + S) of columns [w0, w0 + RW) of image img
  __device__ __forceinline__ bool next(int& img, int& w0, int& hb, int& S) {
    if (u >= u1) return false;
    const int strip = (int)(u / H);
    hb = (int)(u - (long long)strip * H);
    const long long left = u1 - u;
    S = (H - hb) < left ? (H - hb) : (int)left;
    img = strip / strips_w;
    w0 = (strip - img * strips_w) * RW;
    u += S;
    return true;
  }
};

template <int CBLK, int A_STAGES, int OUT_BUFS, int CASEB, int RED = 0>
__global__ void __launch_bounds__(64 + EPI_THREADS, 1) conv3_rows_kernel(const __grid_constant__ Conv3Params p) {
  using L = C3RSmem<CBLK, A_STAGES, OUT_BUFS, CASEB, RED>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* afull = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* aempty = afull + A_STAGES;
  uint64_t* wfull = aempty + A_STAGES;
  uint64_t* rfull = wfull + 1;               // [RING] output row complete (tcgen05.commit)
  uint64_t* rempty = rfull + RING;           // [RING] slot drained and zeroed (EPI_WARPS arrivals)
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
    mbar_init(wfull, 1);
    for (int s = 0; s < RING; ++s) {
      mbar_init(&rfull[s], 1);
      mbar_init(&rempty[s], EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    // resident weights: for every (channel block, column offset) the three filter rows in the order dh = 2, 1, 0 (D column groups 0, 1, 2
    // = output rows y - 1, y, y + 1 of input row y)
    mbar_expect_tx_e(wfull, 9 * CBLK * WT_BYTES);
    for (int cbg = 0; cbg < CBLK; ++cbg)
      for (int dw = 0; dw < 3; ++dw)
        for (int g = 0; g < 3; ++g) {
          const int tap = (2 - g) * 3 + dw;
          tma_load_2d_e(smem + L::OFF_W + ((cbg * 3 + dw) * 3 + g) * WT_BYTES, &p.b_map, wfull, (tap * CBLK + cbg) * 64, 0);
        }
    int stage = 0;
    uint32_t phase = 0;
    RowWalk walk(p);
    int img, w0, hb, S;
    while (walk.next(img, w0, hb, S)) {
      for (int i = -1; i <= S; ++i) {                   // halo row above, S rows, halo row below (out of the image = zero fill)
        for (int src = 0; src < p.nsrc; ++src)
          for (int cb = 0; cb < p.cblk[src]; ++cb) {
            mbar_wait(&aempty[stage], phase ^ 1);
            mbar_expect_tx_e(&afull[stage], ROW_BYTES);
            tma_load_4d_e(smem + L::OFF_A + stage * ROW_STRIDE, &p.a_map[src], &afull[stage], cb * 64, w0 - 1, hb + i, img);
            if (++stage == A_STAGES) { stage = 0; phase ^= 1; }
          }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t tmem_u = warp_uniform(tmem_base);
    const uint32_t smem_base_u = warp_uniform(smem_u32(smem));
    const uint64_t wdesc0 = make_smem_desc(smem_base_u + L::OFF_W, 16, 1024);
    int stage = 0;
    uint32_t phase = 0;
    int rc = 0;                                         // output rows handed to the ring so far (row r of this CTA -> slot r mod 8)
    mbar_wait(wfull, 0u);
    RowWalk walk(p);
    int img, w0, hb, S;
    while (walk.next(img, w0, hb, S)) {
      for (int i = -1; i <= S; ++i) {
        // input row i feeds the output rows j = i - 1 + g (g = 0, 1, 2) that belong to the segment
        const int g_lo = i < 1 ? 1 - i : 0;
        const int g_hi = S - i < 2 ? S - i : 2;
        if (g_hi == 2) {                                // first touch of output row i + 1: its slot must be drained and zeroed
          const int r = rc + i + 1;
          mbar_wait(&rempty[r & (RING - 1)], (uint32_t)(r >> 3) & 1u);
        }
        const int n = g_hi - g_lo + 1;                  // active 64-column groups
        const int b = (rc + i - 1 + g_lo) & (RING - 1); // slot of the first one
        const int n1 = n < RING - b ? n : RING - b;     // groups before the ring wraps
        const uint32_t idesc1 = make_idesc_bf16(128, 64 * n1, 0, 0);
        const uint32_t idesc2 = make_idesc_bf16(128, 64 * (n - n1 > 0 ? n - n1 : 1), 0, 0);
        for (int cbg = 0; cbg < CBLK; ++cbg) {
          mbar_wait(&afull[stage], phase);
          tc_fence_after();
          const uint64_t adesc0 = make_smem_desc(smem_base_u + L::OFF_A + stage * ROW_STRIDE, 16, 1024);
#pragma unroll
          for (int dw = 0; dw < 3; ++dw) {
            const uint64_t adesc = adesc0 + (uint64_t)((dw * 128) >> 4);
            const uint64_t wdesc = wdesc0 + (uint64_t)((((cbg * 3 + dw) * 3 + g_lo) * WT_BYTES) >> 4);
            tc_mma4_bf16_e(tmem_u + b * 64, adesc, wdesc, idesc1, 1u);
            if (n1 < n) tc_mma4_bf16_e(tmem_u, adesc, wdesc + (uint64_t)((n1 * WT_BYTES) >> 4), idesc2, 1u);
          }
          tc_commit_e(&aempty[stage]);
          if (++stage == A_STAGES) { stage = 0; phase ^= 1; }
        }
        if (i >= 1) tc_commit_e(&rfull[(rc + i - 1) & (RING - 1)]);       // output row i - 1 has all three input rows
      }
      rc += S;
    }
    __syncwarp();
  } else {
    // ================= epilogue =================
    Epilogue<64, OUT_BUFS, RW, RED, CASEB> epi(smem + L::OFF_EPI, p.ep, tmem_base, nullptr, nullptr, threadIdx.x - 64, warp);
    epi.red_map = &p.red_map;
    epi.load_vectors(0);
    // all 512 accumulator columns start at zero: this warp's 32 lanes x its half of the columns
    const uint32_t lane_addr = tmem_base + ((uint32_t)(epi.quad * 32) << 16);
#pragma unroll
    for (int c = 0; c < 8; ++c) tmem_st_zero_32x32(lane_addr + epi.half * 256 + c * 32);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0)
      for (int s = 0; s < RING; ++s) mbar_arrive(&rempty[s]);
    int rc = 0;
    RowWalk walk(p);
    int img, w0, hb, S;
    bool more = walk.next(img, w0, hb, S);
    while (more) {
      RowWalk peek = walk;                                // the segment after this one (RED == 2 prefetches the next row's `a` tile)
      int nimg = 0, nw0 = 0, nhb = 0, nS = 0;
      const bool more_next = peek.next(nimg, nw0, nhb, nS);
      for (int j = 0; j < S; ++j) {
        const int r = rc + j;
        const int slot = r & (RING - 1);
        mbar_wait(&rfull[slot], (uint32_t)(r >> 3) & 1u);
        const int h = hb + j;
        const bool in_seg = j + 1 < S;
        epi.tile(h, w0, [&](const uint8_t* blk, int) { tma_store_4d(&p.o_map[0], blk, 0, w0, h, img); }, slot * 64, false, false, 0, img,
                 in_seg || more_next, in_seg ? h + 1 : nhb, in_seg ? w0 : nw0, in_seg ? img : nimg);
        // the accumulator has been read (tcgen05.wait::ld inside tile): zero it and hand the slot to output row r + 8
        tmem_st_zero_32x32(lane_addr + slot * 64 + epi.half * 32);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&rempty[slot]);
      }
      rc += S;
      walk = peek;
      img = nimg; w0 = nw0; hb = nhb; S = nS;
      more = more_next;
    }
    epi.finish(0, blockIdx.x);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// host side: p.o_base / p.o_ch, p.H, p.W, p.ep and p.fin are set by the entry points below
template <int CBLK, int A_STAGES, int OUT_BUFS, int CASEB, int RED = 0>
int launch_c3_rows(Conv3Params& p, const void* const* a_base, const int* a_ch, int n_img, cudaStream_t stream) {
  using L = C3RSmem<CBLK, A_STAGES, OUT_BUFS, CASEB, RED>;
  static_assert(L::TOTAL <= 232448, "smem budget");
  auto kern = conv3_rows_kernel<CBLK, A_STAGES, OUT_BUFS, CASEB, RED>;
  static bool attr_done = false;
  if (!attr_done) {
    UB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_done = true;
  }
  int rc;
  for (int i = 0; i < p.nsrc; ++i)
    if ((rc = ub_tmap_act4d(&p.a_map[i], a_base[i], a_ch[i], p.W, p.H, n_img, (long long)a_ch[i] * 2, (long long)p.W * a_ch[i] * 2,
                            (long long)p.H * p.W * a_ch[i] * 2, RPW, 1)))
      return rc;
  if ((rc = ub_tmap_mat2d(&p.b_map, p.w_base, 64, 9ll * CBLK * 64, 64))) return rc;
  if ((rc = ub_tmap_act4d(&p.o_map[0], p.o_base[0], p.o_ch[0], p.W, p.H, n_img, (long long)p.o_ch[0] * 2, (long long)p.W * p.o_ch[0] * 2,
                          (long long)p.H * p.W * p.o_ch[0] * 2, RW, 1)))
    return rc;
  if (RED && (rc = ub_tmap_act4d(&p.red_map, p.red_base, p.ep.red_ncols, p.W, p.H, n_img, (long long)p.ep.red_ncols * 2,
                                 (long long)p.W * p.ep.red_ncols * 2, (long long)p.H * p.W * p.ep.red_ncols * 2, RW, 1)))
    return rc;
  p.strips_w = (p.W + RW - 1) / RW;         // a ragged last strip: loads zero-fill, stores clip, statistics mask (epilogue.cuh)
  const long long total = (long long)n_img * p.strips_w * p.H;
  UB_CHECK_SHAPE(total > 0 && total < (1ll << 30), "conv3 (rows): row count out of range");
  p.rows_total = (int)total;
  p.n_tiles = 1;
  long long grid = ub_num_sms();
  if (grid > total) grid = total;
  UB_CHECK_SHAPE(grid <= UB_STATS_ROWS, "conv3 (rows): stats rows");
  const BnFin* fin = p.fin;
  p.fin = nullptr;
  const bool fused = fin && p.ep.stats && epi_fin_bytes(p.ncols) <= OUT_BUFS * L::E::OUT_BYTES;
  if (fused) epi_set_fin(p.ep, *fin, (int)grid);
  else if (p.ep.stats) UB_CUDA(cudaMemsetAsync(p.ep.stats, 0, sizeof(float) * UB_STATS_ROWS * 2 * p.ncols, stream));
  if (RED) UB_CUDA(cudaMemsetAsync(p.ep.red_out, 0, sizeof(float) * UB_STATS_ROWS * 2 * p.ep.red_ncols, stream));
  kern<<<(int)grid, 64 + EPI_THREADS, L::TOTAL, stream>>>(p);
  UB_LAUNCH_CHECK();
  if (fin && !fused)
    return ub_bn_finalize(p.ep.stats, p.ncols, fin->groups, (long long)fin->count, fin->mean, fin->rstd, fin->moving_mean, fin->moving_var,
                          fin->momentum, fin->eps, stream);
  return UB_OK;
}

// the last strip of a row may be ragged; worth it while the padding stays under 1/8 of the row
static bool rows_width_ok(int W) { return W >= RW && ((W + RW - 1) / RW * RW - W) * 8 <= W; }

// UB_CONV3_ROWS: 0 = off (the halo-patch kernels of igemm_conv3.cu), 1 = 64 -> 64 layers only, 2 = also 128 -> 64 (dec1a forward, enc2a
// dgrad), 3 (default) = also the 64 -> 64 + 64 dgrad of dec1a as two launches (igemm_conv3.cu: launch()).  Measured on B200
// (profiles/r02_ab_runs.md, blocks N-R), sustained at the 1000 W cap: enc1b forward 707 -> 968 TFLOP/s, enc1b dgrad 796 -> 1021, dec1a
// forward 908 -> 1158, dec1a dgrad 930 -> 1084, enc2a dgrad 905 -> 1137; the step 22.7 -> 20.9 ms together with the epilogue changes.
static int use_rows() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("UB_CONV3_ROWS");
    v = e ? atoi(e) : 3;
  }
  return v;
}
