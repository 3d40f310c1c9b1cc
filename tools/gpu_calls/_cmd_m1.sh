export PYTHONPATH=.
for i in 1 2; do
timeout 600 python bench.py --workload c5 --steps 3 > gpurun_out/m1_bench_c5_$i.json 2> gpurun_out/m1_bench_c5_$i.err
python -c "
import json
d=json.loads(open('gpurun_out/m1_bench_c5_$i.json').read().strip().splitlines()[-1])
print('c5 ms_per_step %.4f e2e %.3f cold %s frac %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_cold_pageable'], d['roofline']['frac']))
"
done
