export B200CLIP_ALLOW_SYNTHETIC=1
timeout 300 python -m pytest tests/test_gpu_preprocess.py -x -q -m gpu -k sm_reserve 2>&1 | tail -2
for v in "--sm-reserve 0" "--sm-reserve 16" "--sm-reserve 24" "--no-text-overlap" "--sm-reserve 8" "--sm-reserve 16" "--sm-reserve 0"; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e $v > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
python - <<P
import json
for l in open('gpurun_out/bench_ab.json'):
    if l.startswith('{'):
        d=json.loads(l); print('[$v]', round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k in ('gemm','attention','pre_area','pre_vpass','head','gemm_small','misc')}, d['clocks']['sm_mhz'])
P
done
