set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_batcher.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -4
K2_TMA_XYZ="0,1" K2_VARIANTS="xyz_only,xyz_rd,all" timeout 900 python tools/k2_parts.py 2>&1 | tail -8
