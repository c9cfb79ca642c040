tools/gpu_round.sh r2A lookup_kernel:3:full extract_identify_kernel:3:full score_dense_kernel:3:full
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2A_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2A_launches.log 2>&1
