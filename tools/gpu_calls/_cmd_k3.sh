export PYTHONPATH=.
timeout 600 python -m pytest tests/test_rig_gpu.py tests/test_pycalibrator_gpu.py -x -q 2>&1 | tail -5
timeout 600 python -m pytest tests/test_full_size_gpu.py -x -q -k "c3" 2>&1 | tail -3
timeout 300 python tools/profile_rig.py > gpurun_out/k3_rig_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rig_sweep_tc -s 3 -c 1 -f -o gpurun_out/k3_rig_sweep_tc python tools/profile_rig.py > gpurun_out/k3_rig_ncu.log 2>&1
echo "ncu exit $?"
bash tools/gpu_scale.sh k3 2 c3 2>&1 | grep -E "exit"
python -c "
import json
d=json.loads(open('gpurun_out/k3_c3_n2.json').read().strip().splitlines()[-1])
print('c3 n2 ms', d['ms_per_step'], d.get('partition_check'))
"
