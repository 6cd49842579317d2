/* Hardware probes: NOT part of libunetb200.so.  `make -C semantic-segmentation-unet_b200/csrc probes` builds them into
 * libunetb200_probe.so for tools/desc_probe.py and tools/mma_rate.py, which measured the swizzled-operand addressing and the MMA issue
 * rates the conv kernels rely on (profiles/r01g_mma_rate.jsonl). */
#ifndef UNETB200_PROBE_H_
#define UNETB200_PROBE_H_
#include <cuda_runtime.h>
#ifdef __cplusplus
extern "C" {
#endif
int ub_debug_mma_rate(int N, int iters, int mn_major, int a_shift_rows, int a_sbo_bytes, int nblocks, long long* clocks,
                      cudaStream_t stream);
int ub_debug_desc_probe(const void* x, int R, const void* ident, float* out, int shift, int sbo_bytes, int base_offset, int mode,
                        cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* UNETB200_PROBE_H_ */
