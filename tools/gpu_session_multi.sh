#!/bin/bash
# Multi-GPU session (run under `gpurun --gpus N`): numerical multi-GPU tests + the bench line at N ranks.
set -u
N=${1:-2}
OUT=gpurun_out/multi$N
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,power.limit --format=csv > $OUT/gpu.txt 2>&1
timeout 400 python -m pytest tests/test_multigpu.py -m gpu -q > $OUT/pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_multigpu.log
cp gpurun_out/parity_measured.jsonl $OUT/ 2>/dev/null
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
tail -5 $OUT/pytest_multigpu.log
head -c 400 $OUT/bench.json
tail -3 $OUT/bench.err
