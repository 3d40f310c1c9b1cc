#!/bin/bash
# per-launch device times of one solve (ncu launch list; cold-cache, serialised: compare shares).
# Usage: tools/gpu_launches.sh <tag> <views> <points> [env assignments...]
tag=$1; nv=$2; np=$3; shift 3
mkdir -p gpurun_out
export PYTHONPATH=.
for e in "$@"; do export "$e"; done
timeout 300 python tools/profile_sweep.py $nv $np > gpurun_out/${tag}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_sweep.py $nv $np > gpurun_out/${tag}_ncu.log 2>&1
echo "exit $?"
