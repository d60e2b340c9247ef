export B200CLIP_ALLOW_SYNTHETIC=1
B200CLIP_K1_VERBOSE=1 timeout 300 python tests/k1_variant_check.py 2>&1 | tail -2
timeout 300 python tools/bench_k1.py 1024 2>&1 | tail -1
timeout 300 python tools/bench_k1.py 1024 720 1280 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_mma.json 2> gpurun_out/bench_mma.err
python - <<'P'
import json
for f in ('bench_mma',):
    for l in open(f'gpurun_out/{f}.json'):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value']), d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k in ('gemm','attention','pre_area','pre_vpass','preprocess')}, d['clocks']['sm_mhz'])
P
