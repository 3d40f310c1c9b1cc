export PYTHONPATH=.
timeout 600 python -m pytest tests/test_rig_gpu.py tests/test_pycalibrator_gpu.py -x -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_full_size_gpu.py -x -q -k "c3" 2>&1 | tail -2
for w in c3; do
timeout 300 python bench.py --workload $w --steps 5 > gpurun_out/i8_bench_$w.json 2> gpurun_out/i8_bench_$w.err; tail -c 300 gpurun_out/i8_bench_$w.err
python -c "
import json
d=json.loads(open('gpurun_out/i8_bench_$w.json').read().strip().splitlines()[-1])
print('$w ms_per_step', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'), d['solve'])
"
done
timeout 300 python tools/profile_rig.py > gpurun_out/i8_rig_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/i8_rig_launches.csv python tools/profile_rig.py > gpurun_out/i8_rig_ncu.log 2>&1
tail -1 gpurun_out/i8_rig_plain.log | cut -c1-330
timeout 600 python tools/e2e_breakdown.py 2>&1 | tail -12
