export PYTHONPATH=.
timeout 900 python -m pytest tests/test_intrinsics_gpu.py tests/test_pycalibrator_gpu.py tests/test_full_size_gpu.py -x -q 2>&1 | tail -2
for e in 0 1; do
  if [ $e = 1 ]; then export CCG_NO_PDL=1; else unset CCG_NO_PDL; fi
  echo "== CCG_NO_PDL=$e"
  for w in c5 c4 c2; do
    timeout 300 python bench.py --workload $w --steps 5 > gpurun_out/j2_bench_${w}_$e.json 2> gpurun_out/j2_bench_${w}_$e.err
    python -c "
import json
d=json.loads(open('gpurun_out/j2_bench_${w}_$e.json').read().strip().splitlines()[-1])
print('$w ms_per_step %.4f e2e %.3f launch_ms %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], (d.get('roofline') or {}).get('launch_ms')))
"
  done
  bash tools/gpu_scale.sh j2_$e 2 c5 2>&1 | grep -E "exit"
  python -c "
import json
d=json.loads(open('gpurun_out/j2_${e}_c5_n2.json').read().strip().splitlines()[-1])
print('c5 n2 ms', d['ms_per_step'], d.get('partition_check'), d.get('breakdown_ms_per_step'))
"
done
