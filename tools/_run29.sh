set -x
mkdir -p gpurun_out
MAXH=64 timeout 900 python tools/sweep_k2_ctas.py 2>&1 | tail -34 | cut -c1-150
OB_DECODE_PIPE_CTAS=3 timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_batcher.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_batcher.py tests/test_gpu_pipeline.py tests/test_gpu_python_api.py -x -q -m gpu 2>&1 | tail -3
