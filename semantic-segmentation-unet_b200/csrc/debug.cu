// Hardware probe (test infrastructure of the library itself, not on the product path): how does tcgen05.mma address a
// SWIZZLE_128B shared-memory operand whose descriptor START ADDRESS is a multiple of 128 B but not of the 1024-byte
// swizzle atom, and whose 8-row group stride (SBO) is not a multiple of 1024 B?  The implicit-GEMM kernels want to read
// the 9 filter taps of a 3x3 conv as shifted views of ONE halo'd activation patch that TMA wrote once; that only
// works if the XOR pattern is a function of the absolute shared-memory address (as it is for TMA writes).
//
//   X  [R][64] bf16 (global)  --TMA, SWIZZLE_128B-->  smem rows r at 128-byte pitch (1024-aligned base)
//   mode 0 (K-major A):  D[m][n] = sum_k A[m][k] I[n][k],  row(m) = shift + (m / 8) * (sbo / 128) + m % 8   -> D[m][n] = X[row(m)][n]
//   mode 1 (MN-major A): A[m = channel][k = pixel]:        D[c][n] = X[shift + (n / 8) * (sbo / 128) + n % 8][c]   (c < 64; rows 64..127 repeat)
#include "common.cuh"

namespace {

struct ProbeParams {
  CUtensorMap x_map;   // [R][64]
  CUtensorMap i_map;   // identity [64][64]
  int R;
  int shift, sbo, base_offset, mode;
  float* out;          // [128][64]
};

__global__ void __launch_bounds__(128, 1) desc_probe_kernel(const __grid_constant__ ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sx = smem;                    // up to 256 rows x 128 B = 32 KB
  uint8_t* si = smem + 32768;            // 64 x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 32768 + 8192);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], p.R * 128 + 64 * 128);
    tma_load_2d(sx, &p.x_map, &bars[0], 0, 0);
    tma_load_2d(si, &p.i_map, &bars[0], 0, 0);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t ax = smem_u32(sx) + p.shift * 128;
    const uint32_t bx = smem_u32(si);
    const uint64_t bo = (uint64_t)(p.base_offset & 7) << 49;
    if (p.mode == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        tc_mma_bf16(tmem_base, (make_smem_desc(ax, 16, p.sbo) | bo) + 2 * k, make_smem_desc(bx, 16, 1024) + 2 * k, idesc, k != 0);
    } else {
      const uint32_t idesc = make_idesc_bf16(128, 64, 1, 0);
      const int rows_per_group = p.sbo / 128;
#pragma unroll
      for (int k = 0; k < 4; ++k)      // K = 16 pixels per MMA = 2 groups of 8 rows
        tc_mma_bf16(tmem_base, make_smem_desc(ax + k * 2 * rows_per_group * 128, 0, p.sbo) | bo, make_smem_desc(bx, 16, 1024) + 2 * k, idesc, k != 0);
    }
    tc_commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
#pragma unroll
  for (int chunk = 0; chunk < 2; ++chunk) {
    uint32_t v[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + chunk * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) p.out[row * 64 + chunk * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

// Issue-rate probe: one CTA per SM issues `iters` x 4 MMAs (M = 128, K = 16 each) on operands that stay resident in shared
// memory (contents irrelevant) and reports elapsed SM clocks: the tensor-pipe + smem-read ceiling per tile shape.
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int iters, int mn_major, int a_shift, int a_sbo, long long* clocks) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (49152 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 256);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t sa = smem_u32(smem) + a_shift * 128, sb = smem_u32(smem) + 49152;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)mn_major << 15) | ((uint32_t)mn_major << 16) |
                           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = mn_major ? make_smem_desc(sa + k * 2048, 8192, a_sbo) : make_smem_desc(sa, 16, a_sbo) + 2 * k;
        const uint64_t bd = mn_major ? make_smem_desc(sb + k * 2048, 8192, 1024) : make_smem_desc(sb, 16, 1024) + 2 * k;
        tc_mma_bf16(tmem_base, ad, bd, idesc, 1);
      }
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    clocks[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

extern "C" int ub_debug_mma_rate(int N, int iters, int mn_major, int a_shift, int a_sbo, int nblocks, long long* clocks, cudaStream_t stream) {
  UB_CHECK_ARG(clocks && N >= 16 && N <= 256 && N % 16 == 0 && iters > 0 && nblocks > 0, "mma_rate: bad args");
  const int smem = 49152 + 32768 + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    UB_CUDA(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  mma_rate_kernel<<<nblocks, 128, smem, stream>>>(N, iters, mn_major, a_shift, a_sbo, clocks);
  UB_LAUNCH_CHECK();
  return UB_OK;
}

extern "C" int ub_debug_desc_probe(const void* x, int R, const void* ident, float* out, int shift, int sbo_bytes, int base_offset, int mode,
                                   cudaStream_t stream) {
  UB_CHECK_ARG(x && ident && out && R >= 8 && R <= 256 && shift >= 0 && sbo_bytes % 128 == 0, "desc_probe: bad args");
  ProbeParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = ub_tmap_mat2d(&p.x_map, x, R, 64, R))) return rc;
  if ((rc = ub_tmap_mat2d(&p.i_map, ident, 64, 64, 64))) return rc;
  p.R = R;
  p.shift = shift;
  p.sbo = sbo_bytes;
  p.base_offset = base_offset;
  p.mode = mode;
  p.out = out;
  const int smem = 32768 + 8192 + 64 + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    UB_CUDA(cudaFuncSetAttribute(desc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  desc_probe_kernel<<<1, 128, smem, stream>>>(p);
  UB_LAUNCH_CHECK();
  return UB_OK;
}
