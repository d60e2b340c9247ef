export B200CLIP_ALLOW_SYNTHETIC=1
for c in 2 3; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --config $c --no-cpu > gpurun_out/bench_n2_cfg$c.json 2> gpurun_out/bench_n2_cfg$c.err
python - <<P
import json
for l in open('gpurun_out/bench_n2_cfg$c.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=2 cfg$c', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d.get('per_rank_ms_per_step'), d.get('per_rank_allgather_ms_per_step'), d['clocks']['sm_mhz'])
P
done
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -2
