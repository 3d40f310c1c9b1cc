#!/bin/bash
# Sweep timing of variant builds only.  Usage: tools/gpu_time.sh <tag> [variants...]
tag=$1; shift
mkdir -p gpurun_out
export PYTHONPATH=.
: > gpurun_out/${tag}_times.log
for v in "$@"; do
  CCG_LIB=$PWD/camera_calibrator_b200/variants/libccg_$v.so timeout 300 python tools/time_sweep.py 100000 400 2>&1 | grep -v "^ -\|^   solve" >> gpurun_out/${tag}_times.log
  echo "exit $? ($v)" >> gpurun_out/${tag}_times.log
done
cat gpurun_out/${tag}_times.log
