export PYTHONPATH=.
bash tools/gpu_scale.sh s8b 8 c5 c3 c4 2>&1 | grep -E "exit"
python - <<'P'
import json
for w in ['c5','c3','c4']:
    try:
        d=json.loads(open(f'gpurun_out/s8b_{w}_n8.json').read().strip().splitlines()[-1])
        print(w, 'n8 ms', d['ms_per_step'], 'value %.4g'%d['value'], d.get('partition_check'), d.get('breakdown_ms_per_step'))
    except Exception as e:
        print(w, 'ERR', e)
P
