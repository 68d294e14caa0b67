set -x
mkdir -p gpurun_out
K2_SPLITS="1,1;2,1;4,1;8,1;16,1;8,2;8,4;1,4;8,8" K2_VARIANTS="none,xyz_only,all" timeout 900 python tools/k2_parts.py 2>&1 | tail -30
timeout 600 python -m pytest tests/test_gpu_decode.py -x -q -m gpu 2>&1 | tail -3
