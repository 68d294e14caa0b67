set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_python_api.py tests/test_gpu_batcher.py -x -q -m gpu 2>&1 | tail -15
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_pipe -s 4 -c 1 -o gpurun_out/k2_pipe_r02 -f python bench.py --only k2 --kernel-only --steps 3 --warmup 3 > gpurun_out/ncu_k2.log 2>&1
tail -3 gpurun_out/ncu_k2.log
