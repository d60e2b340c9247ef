set -x
export B200CLIP_ALLOW_SYNTHETIC=1
timeout 900 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_nv12.py -x -q -m gpu > gpurun_out/k1_tests.log 2>&1; tail -3 gpurun_out/k1_tests.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_mma.json 2> gpurun_out/bench_mma.err
B200CLIP_AREA_NOMMA=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_nomma.json 2> gpurun_out/bench_nomma.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_mma2.json 2> gpurun_out/bench_mma2.err
python - <<'P'
import json
for f in ('bench_mma','bench_nomma','bench_mma2'):
    for l in open(f'gpurun_out/{f}.json'):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['value']), d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k in ('gemm','attention','pre_area','pre_vpass','preprocess')}, d['clocks']['sm_mhz'], d['e2e']['value'])
P
