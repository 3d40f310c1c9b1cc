export PYTHONPATH=.
timeout 600 python -m pytest tests/test_rig_gpu.py tests/test_pycalibrator_gpu.py -x -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_full_size_gpu.py -x -q -k c3 2>&1 | tail -2
timeout 300 python bench.py --workload c3 --steps 5 > gpurun_out/i5_bench_c3.json 2> gpurun_out/i5_bench_c3.err; tail -c 300 gpurun_out/i5_bench_c3.err
python -c "
import json
d=json.loads(open('gpurun_out/i5_bench_c3.json').read().strip().splitlines()[-1])
print('c3 ms_per_step', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'), d['solve'])
"
timeout 300 python tools/profile_rig.py > gpurun_out/i5_rig_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/i5_rig_launches.csv python tools/profile_rig.py > gpurun_out/i5_rig_ncu.log 2>&1
tail -1 gpurun_out/i5_rig_plain.log | cut -c1-330
