set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_edge_cases.py tests/test_gpu_batcher.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -15
for pipe in 1 0; do
  OB_DECODE_PIPE=$pipe timeout 300 python bench.py --workload k2 --kernel-only --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/k2_pipe$pipe.json
done
OB_DECODE_PIPE=1 OB_DECODE_PIPE_WARPS=12 timeout 300 python bench.py --workload k2 --kernel-only --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/k2_pipe1_w12.json
OB_DECODE_PIPE=1 timeout 300 python bench.py --workload k2 --kernel-only --steps 20 --warmup 5 --streams-per-gpu 8 2>&1 | tail -1 | tee gpurun_out/k2_pipe1_s8.json
OB_DECODE_PIPE=0 timeout 300 python bench.py --workload k2 --kernel-only --steps 20 --warmup 5 --streams-per-gpu 8 2>&1 | tail -1 | tee gpurun_out/k2_pipe0_s8.json
