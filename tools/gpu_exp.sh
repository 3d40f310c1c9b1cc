#!/bin/bash
# One GPU session of sweep-kernel experiments: parity tests on the in-tree library, then the sweep
# timing of every variant under camera_calibrator_b200/variants/.  Usage: tools/gpu_exp.sh <tag> [variants...]
tag=$1; shift
mkdir -p gpurun_out
export PYTHONPATH=.
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.log 2>&1
timeout 900 python -m pytest tests/test_intrinsics_gpu.py -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
: > gpurun_out/${tag}_times.log
for v in "$@"; do
  CCG_LIB=$PWD/camera_calibrator_b200/variants/libccg_$v.so timeout 300 python tools/time_sweep.py 100000 400 >> gpurun_out/${tag}_times.log 2>&1
  echo "exit $? ($v)" >> gpurun_out/${tag}_times.log
done
cat gpurun_out/${tag}_times.log
