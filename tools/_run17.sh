set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dewarp.py tests/test_gpu_cloud.py tests/test_gpu_edge_cases.py -x -q -m gpu 2>&1 | tail -8
timeout 600 python tools/time_pose.py > gpurun_out/time_pose.log 2>&1; tail -5 gpurun_out/time_pose.log
