set -x
cd $GRAFT_REPO_ROOT
for cfg in 3 4; do
  FOURIER_B200_CFG=$cfg timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "large_sizes or properties_at_baseline or fused_paths" 2>&1 | tail -3
done > gpurun_out/t50.log 2>&1
for cfg in 0 3 4; do
  echo "== c2 cfg $cfg"; FOURIER_B200_CFG=$cfg timeout 300 python bench.py --batch 1024 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1
done > gpurun_out/b50.log 2>&1
for rl in "16 5" "16 8" "32 8"; do set -- $rl
  echo "== c2 cfg 4 ring $1 lag $2"; FOURIER_B200_CFG=4 FOURIER_B200_RING=$1 FOURIER_B200_LAG=$2 timeout 300 python bench.py --batch 1024 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1
done >> gpurun_out/b50.log 2>&1
for cfg in 0 3 4; do
  echo "== c3 cfg $cfg"; FOURIER_B200_CFG=$cfg timeout 300 python bench.py --workload c3 --batch 16384 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1
done >> gpurun_out/b50.log 2>&1
