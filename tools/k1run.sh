export B200CLIP_ALLOW_SYNTHETIC=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_bench1024.csv python bench.py --frames 1024 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
tail -c 300 gpurun_out/ncu_bench.log
