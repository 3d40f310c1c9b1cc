#!/bin/bash
# Sweep timing of variant builds at several problem sizes.  Usage: tools/gpu_time2.sh <tag> "<views> ..." <points> [variants...]
tag=$1; sizes=$2; pts=$3; shift 3
mkdir -p gpurun_out
export PYTHONPATH=.
: > gpurun_out/${tag}_times.log
for v in "$@"; do
  for nv in $sizes; do
    CCG_LIB=$PWD/camera_calibrator_b200/variants/libccg_$v.so timeout 300 python tools/time_sweep.py $nv $pts 2>&1 | grep -v "^ -" | cut -c1-200 >> gpurun_out/${tag}_times.log
  done
done
cat gpurun_out/${tag}_times.log
