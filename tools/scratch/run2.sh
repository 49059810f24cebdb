cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_twopass -c 1 -o gpurun_out/r01_fused_f64_full -f python bench.py --workload c3 --batch 4096 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --verify 0 > gpurun_out/ncu_c3.log 2>&1
tail -3 gpurun_out/ncu_c3.log
