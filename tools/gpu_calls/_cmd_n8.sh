export PYTHONPATH=.
bash tools/gpu_scale.sh n8 8 c3 c5 2>&1 | grep -E "exit"
python - <<'P'
import json
for w in ['c3','c5']:
    try:
        d=json.loads(open(f'gpurun_out/n8_{w}_n8.json').read().strip().splitlines()[-1])
        print(w, 'n8 ms', d['ms_per_step'], 'value %.4g'%d['value'], d.get('partition_check'), d.get('breakdown_ms_per_step'))
    except Exception as e:
        print(w, 'ERR', e)
P
