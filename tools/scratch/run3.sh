cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_parity.py -x -q -k "distributed or pack or six_step or transpose or alternative_persistent" > gpurun_out/t53.log 2>&1
tail -5 gpurun_out/t53.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
PYTHONPATH=. timeout 300 $TR tools/dist_check.py 11 11 > gpurun_out/dist53.log 2>&1
PYTHONPATH=. timeout 300 $TR tools/dist_check.py 13 12 >> gpurun_out/dist53.log 2>&1
grep -i "dist check\|err" gpurun_out/dist53.log | tail
timeout 300 $TR bench.py --gpus 2 --workload c5 --log2n 28 --steps 10 --warmup 3 > gpurun_out/bench_c5_2gpu_53.log 2>&1
tail -1 gpurun_out/bench_c5_2gpu_53.log | cut -c1-400
