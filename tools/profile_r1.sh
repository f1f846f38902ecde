#!/bin/bash
# Round-1 profiling recipe (run under gpurun, one GPU): launch list of the timed region of bench.py + one full capture
# of the dominant kernel.  Outputs land in gpurun_out/; tools/summarise_profile.py turns them into profiles/.
set -x
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches.csv \
    python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r1_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:walk_kernel --launch-skip 3 --launch-count 1 \
    -o gpurun_out/r1_walk_kernel -f python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r1_full.log 2>&1
tail -n 2 gpurun_out/r1_launches_bench.log; tail -n 2 gpurun_out/r1_full.log
