cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_distributed.py -x -q > gpurun_out/t56.log 2>&1
tail -5 gpurun_out/t56.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
PYTHONPATH=. timeout 300 $TR tools/dist_check.py 11 11 peer > gpurun_out/dist56.log 2>&1
PYTHONPATH=. timeout 300 $TR tools/dist_check.py 13 12 peer >> gpurun_out/dist56.log 2>&1
PYTHONPATH=. timeout 300 $TR tools/dist_check.py 12 12 nccl >> gpurun_out/dist56.log 2>&1
grep -i "dist check\|err\|Traceback" gpurun_out/dist56.log | tail
tail -5 gpurun_out/dist56.log
timeout 300 $TR bench.py --gpus 2 --workload c5 --log2n 28 --steps 10 --warmup 3 > gpurun_out/bench_c5_2gpu_56.log 2>&1
tail -1 gpurun_out/bench_c5_2gpu_56.log | cut -c1-330
timeout 300 $TR bench.py --gpus 2 --workload c5 --log2n 28 --steps 10 --warmup 3 --exchange nccl > gpurun_out/bench_c5_2gpu_56n.log 2>&1
tail -1 gpurun_out/bench_c5_2gpu_56n.log | cut -c1-330
