export PYTHONPATH=.
o=gpurun_out; mkdir -p $o
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/multi_gpu_check.py > $o/n6_mgc.log 2>&1
echo "multi_gpu_check exit $?"; grep -E "^\[|MULTI" $o/n6_mgc.log
bash tools/gpu_scale.sh n6 2 c3 2>&1 | grep -E "exit"
python - <<'P'
import json
for w in ['c3']:
    try:
        d=json.loads(open(f'gpurun_out/n6_{w}_n2.json').read().strip().splitlines()[-1])
        print(w, 'n2 ms', d['ms_per_step'], 'value %.4g'%d['value'], d.get('partition_check'), d.get('breakdown_ms_per_step'))
    except Exception as e:
        print(w, 'ERR', e)
P
tail -5 $o/n6_c3_n2.err
