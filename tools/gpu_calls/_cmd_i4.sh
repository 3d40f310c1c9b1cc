export PYTHONPATH=.
timeout 600 python -m pytest tests/test_rig_gpu.py tests/test_intrinsics_gpu.py -x -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_full_size_gpu.py -x -q -k c3 2>&1 | tail -2
timeout 300 python bench.py --workload c3 --steps 5 > gpurun_out/i4_bench_c3.json 2> gpurun_out/i4_bench_c3.err; tail -c 300 gpurun_out/i4_bench_c3.err
python -c "
import json
d=json.loads(open('gpurun_out/i4_bench_c3.json').read().strip().splitlines()[-1])
print('c3 ms_per_step', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'), d['solve'])
"
timeout 300 python tools/profile_rig.py 2>&1 | tail -1 | cut -c1-330
bash tools/gpu_scale.sh i4 2 c5 c3 c4 2>&1 | grep -E "exit|partition_check" | cut -c1-900
