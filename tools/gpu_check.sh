#!/bin/bash
# Quick GPU session: the intrinsics + rig parity tests on the in-tree library, the sweep timing, short bench lines,
# and the launch list of one rig solve.  Usage: tools/gpu_check.sh <tag> [workloads...]
tag=$1; shift
o=gpurun_out; mkdir -p $o
export PYTHONPATH=.
timeout 900 python -m pytest tests/test_intrinsics_gpu.py tests/test_rig_gpu.py -x -q > $o/${tag}_pytest.log 2>&1
echo "pytest exit $?"; tail -2 $o/${tag}_pytest.log
timeout 300 python tools/time_sweep.py 100000 400 2>&1 | grep -v "^ -" | head -3
for w in "$@"; do
  timeout 600 python bench.py --workload $w --steps 3 > $o/${tag}_bench_$w.json 2> $o/${tag}_bench_$w.err
  python - <<P
import json
d=json.loads(open('$o/${tag}_bench_$w.json').read().strip().splitlines()[-1])
print('$w ms_per_step %.4f value %.4g e2e_ms %.3f frac %s launch_ms %s' % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], (d.get('roofline') or {}).get('frac'), (d.get('roofline') or {}).get('launch_ms')))
print('   breakdown', d.get('breakdown_ms_per_step'))
P
done
timeout 300 python tools/profile_rig.py > $o/${tag}_rig_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $o/${tag}_rig_launches.csv python tools/profile_rig.py > $o/${tag}_rig_ncu.log 2>&1
echo "rig launch list exit $?"; tail -1 $o/${tag}_rig_plain.log | cut -c1-400
