mkdir -p gpurun_out/rprof
for c in 65536 131072 32768; do timeout 200 python tools/profile_render.py $c 2>&1 | tail -1; done
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/rprof/launches_render.csv python tools/profile_render.py 65536 > gpurun_out/rprof/log.txt 2>&1
tail -2 gpurun_out/rprof/log.txt
