set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_batcher.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --only k2 --kernel-only --steps 20 --warmup 5 > gpurun_out/k2_only.json 2> gpurun_out/k2_only.err; tail -c 400 gpurun_out/k2_only.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:decode_pipe -s 4 -c 1 -o gpurun_out/k2_pipe_r02f -f python bench.py --only k2 --kernel-only --steps 3 --warmup 3 > gpurun_out/ncu_k2f.log 2>&1; tail -2 gpurun_out/ncu_k2f.log
