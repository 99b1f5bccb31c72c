#!/bin/bash
# 2-GPU (or N-GPU) A/B of the gradient exchange modes of NeRFTrainer's captured step (run under `gpurun --gpus N`)
set -u
N=${1:-2}
OUT=gpurun_out/ddp_ab$N; mkdir -p $OUT
for mode in overlap merged none; do
  NERF_B200_DDP_EXCHANGE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus $N --steps 60 --warmup 5 --deadline 280 > $OUT/bench_$mode.json 2> $OUT/bench_$mode.err
  echo "$mode rc=$?"
  python - $OUT/bench_$mode.json $mode <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], 'ms/step', round(d['ms_per_step'], 4), 'sustained', round(d['sustained']['ms_per_step'], 4), 'e2e', round(d['e2e']['ms_per_step'], 4))
except Exception as e:
    print(sys.argv[2], 'ERR', e)
P
done
