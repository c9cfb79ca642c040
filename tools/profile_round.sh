#!/bin/bash
# ncu evidence for profiles/: launch list of a bench run + one `--set full` capture per hot kernel (run on the GPU box).
#   usage: tools/profile_round.sh <tag>      outputs under gpurun_out/
# Numbers printed by bench.py under ncu are never bench values.
TAG=${1:-r01}
SMALL="--reads-per-step 1000000 --genome-len 5000000 --no-cpu-baseline --no-e2e"     # 1/6.7 of C2, same queries : index ratio
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_${TAG}.log 2>&1
# launch-skip: the index build runs 12 onesweep passes, every identify step 4; 3 warm-up steps precede the timed one
for spec in onesweep_kernel:24 lookup_kernel:3 per_read_taxon_kernel:3 extract_identify_kernel:3 hist_kernel:4; do
    k=${spec%%:*}; skip=${spec##*:}
    ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip $skip --launch-count 1 -f \
        -o gpurun_out/prof_${TAG}_$k python bench.py --steps 1 --warmup 3 $SMALL > gpurun_out/ncu_${TAG}_$k.log 2>&1
done
ls -la gpurun_out | grep ${TAG}
