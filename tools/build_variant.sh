#!/bin/bash
# Build a named variant of libccg.so with extra -D flags (kernel experiments):
#   tools/build_variant.sh <name> [-DFLAG ...]  ->  camera_calibrator_b200/variants/libccg_<name>.so
# Select it at run time with CCG_LIB=<path> (camera_calibrator_b200/_ccg.py).
set -e
cd "$(dirname "$0")/../camera_calibrator_b200/csrc"
name=$1; shift
out=../variants; mkdir -p $out/obj_$name
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in ctx intr lens rig; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC -Xptxas -v "$@" -c $f.cu -o $out/obj_$name/$f.o 2> $out/obj_$name/$f.ptxas.log || { cat $out/obj_$name/$f.ptxas.log; exit 1; }
done
/usr/local/cuda/bin/nvcc $ARCH -shared -o $out/libccg_$name.so $out/obj_$name/*.o -ldl
grep -A2 "sweep_wsILb0" $out/obj_$name/intr.ptxas.log | grep -E "spill|Used" | tr '\n' ' '; echo " -> $name"
