#!/bin/bash
# Load-strategy A/B of the persistent two-pass kernel on one B200 (FOURIER_B200_CFG: 0 = the size's default,
# 1 = the other strategy: TMA staging <-> direct global loads; FOURIER_B200_FUSED=0 = two-launch tile kernels):
#   gpurun --timeout 1500 -- 'bash tools/measure_variants.sh'
# Round 2's first call ran the round-1 variants 0/5/6/7 with this script (profiles/r02_persistent_kernel_variants.txt);
# the losers were deleted, what remains is the per-size choice between the two load strategies.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
export PYTHONPATH=.
: > gpurun_out/variants_bench.log
for real_sizes in "f32 2^13 2^14 2^15 2^16 2^17 2^18 2^19 2^20" "f64 2^12 2^13 2^14 2^16"; do
  set -- $real_sizes; real=$1; shift
  for cfg in 0 1; do FOURIER_B200_CFG=$cfg python tools/size_table.py $real "$@" >> gpurun_out/variants_bench.log 2>&1; done
  FOURIER_B200_FUSED=0 python tools/size_table.py $real "$@" >> gpurun_out/variants_bench.log 2>&1
done
cat gpurun_out/variants_bench.log
