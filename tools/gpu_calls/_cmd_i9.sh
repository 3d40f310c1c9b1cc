export PYTHONPATH=.
timeout 600 python -m pytest tests/test_rig_gpu.py tests/test_pycalibrator_gpu.py tests/test_intrinsics_gpu.py -x -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_full_size_gpu.py -x -q -k "c3" 2>&1 | tail -2
for w in c3 c1; do
timeout 300 python bench.py --workload $w --steps 5 > gpurun_out/i9_bench_$w.json 2> gpurun_out/i9_bench_$w.err; tail -c 300 gpurun_out/i9_bench_$w.err
python -c "
import json
d=json.loads(open('gpurun_out/i9_bench_$w.json').read().strip().splitlines()[-1])
print('$w ms_per_step', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'), (d.get('roofline') or {}).get('launch_ms'))
"
done
bash tools/gpu_scale.sh i9 2 c5 c3 c4 2>&1 | grep -E "exit" 
python - <<'P'
import json
for w in ['c5','c3','c4']:
    d=json.loads(open(f'gpurun_out/i9_{w}_n2.json').read().strip().splitlines()[-1])
    print(w, 'n2 ms', d['ms_per_step'], 'value %.4g'%d['value'], d.get('partition_check'))
P
