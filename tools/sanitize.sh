#!/bin/bash
# compute-sanitizer (memcheck) over a handful of kernel parity cases: one per kernel family, small shapes.  The tcgen05 / TMA kernels
# access shared memory and TMEM through the async proxy, which memcheck does not see; what it does check here is every global-memory
# access of the epilogues, the elementwise kernels, the head kernels and the host-side launch arguments.
#   tools/sanitize.sh > gpurun_out/sanitize.log
# NOTE (round 2): compute-sanitizer is CLOSED on this GPU pool (every invocation exits 86 with "compute-sanitizer is closed on this pool and
# stays closed"), so this sweep has not run; memory safety rests on the NaN-prefilled outputs, the full-tensor checksums and the ragged /
# 2x2 / odd-width cases of tests/kernel_cases.py.
cd "$(dirname "$0")/.."
for c in conv_fwd_64_64 conv_dgrad_bnred_128_128 conv_wgrad_64_128_ragged deconv_bwd_128_64 bn_pool_bf16 bn_bwd_bf16_512 pool_bwd_bnred_bf16 conv_first_c3 conv_first_tiles_c3 head_k8_weighted head_k20_weighted head_argmax_k20 wgrad_folded_64_64 conv_fwd_bn_2x2 bn_sums_wgrad_64_64 adam zscore; do
  out=$(timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/gpu_probe.py --one $c 2>&1)
  rc=$?
  errs=$(echo "$out" | grep -c "Invalid\|Misaligned\|out of bounds")
  summary=$(echo "$out" | grep "ERROR SUMMARY" | tail -1)
  ok=$(echo "$out" | grep -c '"ok": true')
  echo "$c rc=$rc case_ok=$ok errors=$errs $summary"
done
