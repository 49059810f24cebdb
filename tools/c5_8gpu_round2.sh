#!/bin/bash
# One 8-GPU call: correctness of the fused distributed transform at N = 2^30 against the single-GPU three-pass plan,
# the default bench line (c2 weak-sharded + c3 strong-sharded + c5), and the c5 exchange-mode A/B.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
export PYTHONPATH=.
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
$TR tools/dist_check.py 14 16 fused 2>&1 | grep -i "distributed\|oracle\|DIST CHECK\|error" > gpurun_out/dist_check_fused_${N}gpu.txt
cat gpurun_out/dist_check_fused_${N}gpu.txt
$TR bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/bench_${N}gpu_r02b.err | grep '^{' > gpurun_out/bench_${N}gpu_r02b.json
OUT=gpurun_out/c5_modes_${N}gpu.log; : > $OUT
run() { echo "## $*" >> $OUT; env "${@:2}" $TR bench.py --gpus $N --workload c5 --steps 10 --warmup 3 $1 2>&1 | grep '^{' >> $OUT; }
run "--exchange peer" X=1
run "--exchange fused --transposed-output" X=1
run "--exchange fused" FOURIER_B200_DIST_CHUNK_MB=128
N=$N python - <<'PY'
import json
r = json.load(open("gpurun_out/bench_%sgpu_r02b.json" % __import__("os").environ.get("N", "8")))
print("c2", r["value"], r["ms_per_step"], r["roofline"]["frac"], "e2e", r["e2e"]["value"])
for a in r.get("also", []):
    print(a.get("name"), a.get("value"), a.get("ms_per_step"), a.get("config", {}).get("parallelism", "")[:90])
lab = None
for line in open("gpurun_out/c5_modes_%sgpu.log" % __import__("os").environ.get("N", "8")):
    if line.startswith("## "): lab = line[3:].strip()
    elif line.startswith("{"):
        q = json.loads(line); print("%-60s %8.3f ms" % (lab, q["ms_per_step"]))
PY
