set -x
mkdir -p gpurun_out
timeout 900 python tools/sweep_k2_ctas.py 2>&1 | tail -30
OB_DECODE_PIPE_CTAS=2 timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_batcher.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -4
