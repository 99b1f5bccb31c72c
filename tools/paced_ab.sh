mkdir -p gpurun_out/r02p
timeout 120 python tools/kernel_timing.py 2>&1 | grep -E "^---|mma wait|drain" > gpurun_out/r02p/stall_base.txt
for v in paced8 paced16 paced4; do
  NERF_B200_LIB=nerficg_b200/libnerf_b200.$v.so timeout 120 python tools/kernel_timing.py 2>&1 | grep -E "^---|mma wait|drain|Error|error" > gpurun_out/r02p/stall_$v.txt
done
head -30 gpurun_out/r02p/stall_*.txt
