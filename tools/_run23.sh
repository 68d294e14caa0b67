set -x
mkdir -p gpurun_out
K2_DYN="0,1,2" K2_VARIANTS="none,fields_only,xyz_only,all" timeout 900 python tools/k2_parts.py 2>&1 | tail -14
timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_batcher.py -x -q -m gpu 2>&1 | tail -3
