#!/bin/bash
# bench lines of the small workloads with and without the persistent LM kernel.  Usage: tools/gpu_bench_small.sh <tag>
tag=$1
mkdir -p gpurun_out
for e in 0 1; do
  for w in c1 c2 c4 c5; do
    if [ $e = 1 ]; then export CCG_NO_PERSIST=1; else unset CCG_NO_PERSIST; fi
    timeout 400 python bench.py --workload $w --steps 3 > gpurun_out/${tag}_${w}_np$e.json 2> gpurun_out/${tag}_${w}_np$e.err
    python -c "
import json,sys
d=json.loads(open('gpurun_out/${tag}_${w}_np$e.json').read().strip().splitlines()[-1])
print('$w no_persist=$e ms_per_step %.4f e2e %.3f launches %d full %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], (d.get('full_calibration') or {}).get('wall_ms')))
" 2>&1 | tail -1
  done
done
