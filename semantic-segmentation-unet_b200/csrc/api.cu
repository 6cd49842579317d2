// Error plumbing, version, device queries and TMA descriptor encoding for libunetb200.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void ub_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int ub_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int ub_tmap_act4d(CUtensorMap* out, const void* base, int C, int W, int H, int N, long long stride_w_bytes,
                  long long stride_h_bytes, long long stride_n_bytes, int box_w, int box_h) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    ub_set_error("cuTensorMapEncodeTiled driver entry point unavailable");
    return UB_ERR_CUDA;
  }
  UB_CHECK_SHAPE(C % 64 == 0, "tensor map: channel count %d is not a multiple of 64", C);
  UB_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map: base pointer not 16-byte aligned");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)stride_w_bytes, (cuuint64_t)stride_h_bytes, (cuuint64_t)stride_n_bytes};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ub_set_error("cuTensorMapEncodeTiled(4d C=%d W=%d H=%d N=%d) failed: %d", C, W, H, N, (int)r);
    return UB_ERR_CUDA;
  }
  return UB_OK;
}

int ub_tmap_mat2d(CUtensorMap* out, const void* base, long long rows, long long K, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    ub_set_error("cuTensorMapEncodeTiled driver entry point unavailable");
    return UB_ERR_CUDA;
  }
  UB_CHECK_SHAPE(K % 64 == 0 && box_rows <= 256, "tensor map 2d: K=%lld rows box=%d", K, box_rows);
  UB_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map: base pointer not 16-byte aligned");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ub_set_error("cuTensorMapEncodeTiled(2d rows=%lld K=%lld) failed: %d", rows, K, (int)r);
    return UB_ERR_CUDA;
  }
  return UB_OK;
}

// CRC-32C, slicing-by-8 tables built on first use (host only; checkpoint files are a few hundred MB)
static uint32_t g_crc_tab[8][256];
static bool g_crc_ready = false;
static void crc32c_init() {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
    g_crc_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_crc_tab[t][i] = (g_crc_tab[t - 1][i] >> 8) ^ g_crc_tab[0][g_crc_tab[t - 1][i] & 0xff];
  g_crc_ready = true;
}

extern "C" {
long long ub_host_crc32c(const void* data, long long n, long long crc_in) {
  if (!g_crc_ready) crc32c_init();
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = ~static_cast<uint32_t>(crc_in);
  while (n > 0 && (reinterpret_cast<uintptr_t>(p) & 7)) {
    c = g_crc_tab[0][(c ^ *p++) & 0xff] ^ (c >> 8);
    --n;
  }
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= c;
    c = g_crc_tab[7][w & 0xff] ^ g_crc_tab[6][(w >> 8) & 0xff] ^ g_crc_tab[5][(w >> 16) & 0xff] ^ g_crc_tab[4][(w >> 24) & 0xff] ^
        g_crc_tab[3][(w >> 32) & 0xff] ^ g_crc_tab[2][(w >> 40) & 0xff] ^ g_crc_tab[1][(w >> 48) & 0xff] ^ g_crc_tab[0][w >> 56];
    p += 8;
    n -= 8;
  }
  while (n-- > 0) c = g_crc_tab[0][(c ^ *p++) & 0xff] ^ (c >> 8);
  return static_cast<long long>(~c & 0xffffffffu);
}
const char* ub_last_error(void) { return g_err; }
int ub_version(void) { return UB_VERSION; }
int ub_device_sm_count(void) { return ub_num_sms(); }
}
