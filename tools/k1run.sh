export B200CLIP_ALLOW_SYNTHETIC=1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
ncu --set full --import-source on --clock-control none -k regex:attention_tc64 -s 9 -c 1 -f -o gpurun_out/attn_tc64 python tools/attn_tc64_check.py > gpurun_out/ncu_attn64.log 2>&1; tail -1 gpurun_out/ncu_attn64.log
