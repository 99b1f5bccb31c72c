#!/bin/bash
# cp.async weight-ring variants: stall accounting, and for every variant that runs, the MLP / renderer parity tests and a bench line
set -u
OUT=gpurun_out/${TAG:-lsuw}; mkdir -p $OUT
timeout 150 python tools/kernel_timing.py > $OUT/stall_base.txt 2>&1
for v in ${VARIANTS:-lsuw lsuw1 lsuw2}; do
  NERF_B200_LIB=nerficg_b200/libnerf_b200.$v.so timeout 150 python tools/kernel_timing.py > $OUT/stall_$v.txt 2>&1
  rc=$?; echo "rc=$rc" >> $OUT/stall_$v.txt
  if [ $rc -eq 0 ] && [ -z "${QUICK:-}" ]; then
    timeout 600 python tools/with_variant.py nerficg_b200/libnerf_b200.$v.so pytest tests/test_mlp_gpu.py tests/test_renderer_gpu.py -m gpu -q > $OUT/pytest_$v.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_$v.log
    timeout 400 python tools/with_variant.py nerficg_b200/libnerf_b200.$v.so bench.py --steps 30 --warmup 5 > $OUT/bench_$v.json 2> $OUT/bench_$v.err; echo "rc=$?" >> $OUT/bench_$v.err
  elif [ $rc -ne 0 ]; then
    NERF_B200_LIB=nerficg_b200/libnerf_b200.$v.so timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/kernel_timing.py 512 192 > $OUT/sanitizer_$v.txt 2>&1
  fi
done
grep -H "^---\|^rc=\|wait bias\|slot barrier\|mask" $OUT/stall_*.txt
for f in $OUT/pytest_*.log; do echo $f; tail -3 $f; done 2>/dev/null
for f in $OUT/sanitizer_*.txt; do echo $f; grep -m 12 -A 12 "=========" $f | head -40; done 2>/dev/null
