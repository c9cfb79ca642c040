#!/bin/bash
# compute-sanitizer over the GPU parity tests that reach the round-2 kernels (run on the GPU box; output: gpurun_out/r2_sanitizer.txt)
OUT=gpurun_out/r2_sanitizer.txt
SEL='extract or identify_k12_six or identify_k25 or spills or large_taxonomies or three_shards or long_reads or c1_k12_7'
echo "# memcheck: compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_reference_scale.py -m gpu -k '$SEL'" > $OUT
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_reference_scale.py -m gpu -q -x -k "$SEL" 2>&1 | tail -3 >> $OUT
SEL2='extract or identify_k12_six or identify_k25 or spills'
echo "# racecheck: compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -k '$SEL2'" >> $OUT
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL2" 2>&1 | tail -3 >> $OUT
cat $OUT
