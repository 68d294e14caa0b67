set -x
mkdir -p gpurun_out
K2_HELPERS="0,3,5,6" K2_VARIANTS="fields_only,all" timeout 900 python tools/k2_parts.py 2>&1 | tail -10
OB_DECODE_PIPE_HELPERS=5 timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_batcher.py -x -q -m gpu 2>&1 | tail -3
