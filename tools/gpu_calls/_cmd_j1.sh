export PYTHONPATH=.
timeout 600 python -m pytest tests/test_intrinsics_gpu.py tests/test_pycalibrator_gpu.py -x -q 2>&1 | tail -2
for v in "" late; do
  if [ -n "$v" ]; then export CCG_LIB=$PWD/camera_calibrator_b200/variants/libccg_$v.so; else unset CCG_LIB; fi
  echo "== variant '$v'"
  timeout 300 python tools/time_sweep.py 100000 400 2>&1 | grep "sweep " | head -1
  for w in c4 c2 c1; do
    timeout 300 python bench.py --workload $w --steps 5 > gpurun_out/j1_bench_${w}_$v.json 2> gpurun_out/j1_bench_${w}_$v.err
    python -c "
import json
d=json.loads(open('gpurun_out/j1_bench_${w}_$v.json').read().strip().splitlines()[-1])
print('$w ms_per_step %.4f e2e %.3f launch_ms %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], (d.get('roofline') or {}).get('launch_ms')))
"
  done
done
