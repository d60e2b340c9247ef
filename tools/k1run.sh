set -x
B200CLIP_ALLOW_SYNTHETIC=1 B200CLIP_K1_VERBOSE=1 B200CLIP_AREA_MMA=1 timeout 300 python tests/k1_variant_check.py > gpurun_out/k1mma_check.log 2>&1; echo rc=$? >> gpurun_out/k1mma_check.log
B200CLIP_AREA_MMA=1 timeout 300 python tools/bench_k1.py 1024 > gpurun_out/k1mma_bench.log 2>&1
timeout 300 python tools/bench_k1.py 1024 > gpurun_out/k1old_bench.log 2>&1
tail -5 gpurun_out/k1mma_check.log; tail -2 gpurun_out/k1mma_bench.log; tail -2 gpurun_out/k1old_bench.log
