#!/bin/bash
# One GPU visit: parity tests, a bench run and ncu captures of the named kernels (outputs under gpurun_out/, tag = $1).
#   usage: tools/gpu_round.sh <tag> [kernel-regex:launch-skip[:full] ...]     (":full" = capture at the bench's full C2 size)
TAG=$1; shift
python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_tests.log 2>&1
tail -3 gpurun_out/${TAG}_tests.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python tools/show.py gpurun_out/${TAG}_bench.json
SMALL="--reads-per-step 1000000 --genome-len 5000000 --no-cpu-baseline --no-e2e"     # 1/6.7 of C2, same queries : index ratio
for spec in "$@"; do
    IFS=: read -r k skip full <<< "$spec"
    SIZE=$SMALL; [ -n "$full" ] && SIZE="--no-cpu-baseline --no-e2e"
    ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip $skip --launch-count 1 -f \
        -o gpurun_out/prof_${TAG}_$k python bench.py --steps 1 --warmup 3 $SIZE > gpurun_out/ncu_${TAG}_$k.log 2>&1
done
