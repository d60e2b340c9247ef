export B200CLIP_ALLOW_SYNTHETIC=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n8_cfg2.json 2> gpurun_out/bench_n8_cfg2.err
python - <<P
import json
for l in open('gpurun_out/bench_n8_cfg2.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=8 cfg2', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), [round(x,2) for x in d.get('per_rank_ms_per_step')], [round(x,3) for x in d.get('per_rank_allgather_ms_per_step')], d['clocks']['sm_mhz'], d.get('h2d_ceiling'))
P
