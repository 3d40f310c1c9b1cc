export PYTHONPATH=.
o=gpurun_out
bash tools/gpu_time2.sh n4 "100000 25000 12500" 400 p8 p4 p8
for v in p8 p4; do
  CCG_LIB=$PWD/camera_calibrator_b200/variants/libccg_$v.so timeout 300 python bench.py --workload c4 --steps 3 > $o/n4_c4_$v.json 2> $o/n4_c4_$v.err
  python - <<P
import json
d=json.loads(open('$o/n4_c4_$v.json').read().strip().splitlines()[-1])
print('$v c4 ms_per_step %.4f launch_ms %s frac %s sweeps %s rms %s' % (d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'], d['solve'].get('sweeps'), d['solve'].get('rms_px')))
P
done
