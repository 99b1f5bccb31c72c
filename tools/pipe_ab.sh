mkdir -p gpurun_out/r02n
for v in pipe_nored pipe_noepi pipe_nostore pipe_none; do
  echo "=== $v" >> gpurun_out/r02n/ab.txt
  NERF_B200_LIB=nerficg_b200/libnerf_b200.$v.so timeout 120 python tools/pipe_timing.py 2>&1 | grep -E "^---|mma wait|mma total|role 1 " >> gpurun_out/r02n/ab.txt
done
cat gpurun_out/r02n/ab.txt
