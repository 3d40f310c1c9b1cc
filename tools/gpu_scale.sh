#!/bin/bash
# bench.py at N ranks on one box (torchrun, one rank per GPU).  Usage: tools/gpu_scale.sh <tag> <N> [workloads...]
tag=$1; n=$2; shift 2
mkdir -p gpurun_out
for w in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $n --workload $w --steps 5 --warmup 3 > gpurun_out/${tag}_${w}_n$n.json 2> gpurun_out/${tag}_${w}_n$n.err
  echo "$w exit $?"; tail -c 600 gpurun_out/${tag}_${w}_n$n.json
done
