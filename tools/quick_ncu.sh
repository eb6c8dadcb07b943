#!/bin/bash
# usage: tools/quick_ncu.sh <impl> [extra metrics]  -- per-launch durations of the engine's kernels (ncu, serialised)
IMPL=${1:-3}
EGSPR_IMPL=$IMPL EGSPR_ITERS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/q_launches.csv python tools/prof_layer.py > /dev/null 2>&1
python tools/launch_shares.py gpurun_out/q_launches.csv
