export B200CLIP_ALLOW_SYNTHETIC=1
timeout 200 python tools/attn_tc2_check.py 2>&1 | tail -1
B200CLIP_ATTN_NOTC2=1 timeout 200 python tools/attn_tc2_check.py 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --config 3 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
B200CLIP_ATTN_NOTC2=1 timeout 600 python bench.py --config 3 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_cfg3_old.json 2> gpurun_out/bench_cfg3_old.err
python - <<'P'
import json
for f in ('bench_cfg3','bench_cfg3_old'):
    for l in open(f'gpurun_out/{f}.json'):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value']), d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k in ('gemm','attention','preprocess')}, d['clocks']['sm_mhz'], d.get('check'))
P
