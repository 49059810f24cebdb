#!/bin/bash
# First GPU call of the next round: run the emulator-only variants on hardware and measure them against the default.
#   gpurun --timeout 1500 -- 'bash tools/measure_variants.sh'            (one B200, ~6 min)
# Writes gpurun_out/variants_{tests,bench}.log.  FOURIER_B200_CFG: 0 default, 5 blocked intermediate, 6 blocked +
# direct loads (no staging), 7 blocked + direct loads in pass 2 only.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
FOURIER_B200_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_distributed.py -q \
  -k "experimental or persistent_variant" > gpurun_out/variants_tests.log 2>&1
tail -5 gpurun_out/variants_tests.log
: > gpurun_out/variants_bench.log
for cfg in 0 5 7 6; do
  echo "== c2 cfg $cfg" >> gpurun_out/variants_bench.log
  FOURIER_B200_CFG=$cfg timeout 300 python bench.py --batch 1024 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 >> gpurun_out/variants_bench.log
  echo "== c3 cfg $cfg" >> gpurun_out/variants_bench.log
  FOURIER_B200_CFG=$cfg timeout 300 python bench.py --workload c3 --batch 16384 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 >> gpurun_out/variants_bench.log
done
python - <<'PY'
import json
for l in open('gpurun_out/variants_bench.log'):
    l = l.strip()
    if l.startswith('=='):
        print(l, end='   ')
    elif l.startswith('{'):
        d = json.loads(l)
        print('%.3e samples/s  %.1f %% of measured HBM peak  %.3f ms/step  verify %s' % (
            d['value'], 100 * d['roofline']['frac'], d['ms_per_step'], d.get('verify')))
    elif l:
        print(l[:160])
PY
