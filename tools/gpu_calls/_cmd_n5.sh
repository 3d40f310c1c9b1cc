export PYTHONPATH=.
o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests -x -q -m gpu > $o/n5_pytest.log 2>&1
echo "pytest exit $?"; tail -3 $o/n5_pytest.log
for w in c4 c5; do
  timeout 600 python bench.py --workload $w --steps 3 > $o/n5_bench_$w.json 2> $o/n5_bench_$w.err
  python - <<P
import json
d=json.loads(open('$o/n5_bench_$w.json').read().strip().splitlines()[-1])
print('$w ms_per_step %.4f value %.4g e2e_ms %.3f frac %s launch_ms %s' % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], (d.get('roofline') or {}).get('frac'), (d.get('roofline') or {}).get('launch_ms')))
print('   breakdown', d.get('breakdown_ms_per_step'))
P
done
