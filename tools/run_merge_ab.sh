mkdir -p gpurun_out
for m in 0 1; do
T1K_MERGE_PARTITIONED=$m timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$m bench.py --gpus 2 --steps 1 --warmup 1 --pairs 100000 > gpurun_out/merge_ab_$m.json 2> gpurun_out/merge_ab_$m.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/merge_ab_$m.json')); c=d['config']; p=d['e2e']['phases_ms']
    print('partitioned=$m', c['read_groups'], c['equivalence_classes'], c['em_iterations'], c['assignments'], 'coalesce_ms %.0f e2e %.0f'%(p['ms_coalesce'], d['e2e']['value']))
except Exception as e:
    print('partitioned=$m FAILED', e); print(open('gpurun_out/merge_ab_$m.err').read()[-800:])
PY
done
