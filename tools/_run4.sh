set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_decode.py tests/test_gpu_edge_cases.py tests/test_gpu_batcher.py tests/test_gpu_pipeline.py tests/test_gpu_python_api.py -x -q -m gpu 2>&1 | tail -8
for i in 1 2; do timeout 300 python bench.py --only k2 --kernel-only --steps 20 --warmup 5 2>&1 | tail -1; done
timeout 300 python bench.py --only k2 --kernel-only --steps 20 --warmup 5 --streams-per-gpu 8 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_pipe -s 4 -c 1 -o gpurun_out/k2_pipe_r02b -f python bench.py --only k2 --kernel-only --steps 3 --warmup 3 > gpurun_out/ncu_k2b.log 2>&1
tail -2 gpurun_out/ncu_k2b.log
