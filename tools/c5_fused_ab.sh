#!/bin/bash
# A/B of the distributed transform's exchange modes on N GPUs (gpurun --gpus N -- 'bash tools/c5_fused_ab.sh N LOG2N'):
# "peer" = row FFTs, then one transposing exchange kernel per step; "fused" = the exchanges that follow row FFTs folded
# into the FFTs' last register stage (csrc/dist_kernels.cuh), with and without the two-stream chunk overlap.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
export PYTHONPATH=.
N=${1:-2}; K=${2:-28}; OUT=gpurun_out/c5_fused_ab_${N}gpu.log
: > $OUT
run() {  # label, env..., -- bench args
  echo "## $*" >> $OUT
  env "${@:2}" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --workload c5 --log2n $K --steps 10 --warmup 3 $1 2>&1 | grep '^{' >> $OUT
}
run "--exchange peer" X=1
run "--exchange fused" X=1
run "--exchange fused" FOURIER_B200_DIST_OVERLAP=0
run "--exchange fused" FOURIER_B200_CHUNK_MB=16
run "--exchange fused" FOURIER_B200_CHUNK_MB=64
run "--exchange peer --transposed-output" X=1
run "--exchange fused --transposed-output" X=1
OUT=$OUT python - <<'PY' >> $OUT.summary
import json,sys
rows=[]
lab=None
for line in open(__import__("os").environ["OUT"]):
    if line.startswith("## "): lab=line[3:].strip()
    elif line.startswith("{"):
        r=json.loads(line); rows.append((lab, r["ms_per_step"], r["value"]))
for lab,ms,v in rows: print("%-70s %8.3f ms  %.3e samples/s" % (lab,ms,v))
PY
cat $OUT.summary
