#!/bin/bash
# ncu capture of the sweep kernel on a quarter-size C5 problem.  Usage: tools/gpu_ncu.sh <tag> [variant] [views] [points]
tag=$1; v=$2; nv=${3:-25000}; np=${4:-400}
mkdir -p gpurun_out
export PYTHONPATH=.
if [ -n "$v" ]; then export CCG_LIB=$PWD/camera_calibrator_b200/variants/libccg_$v.so; fi
timeout 300 python tools/profile_sweep.py $nv $np > gpurun_out/${tag}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_intr_sweep -s 2 -c 1 -f -o gpurun_out/${tag} python tools/profile_sweep.py $nv $np > gpurun_out/${tag}_ncu.log 2>&1
echo "exit $?"; tail -3 gpurun_out/${tag}_plain.log; tail -3 gpurun_out/${tag}_ncu.log
