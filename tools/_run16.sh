set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dewarp.py tests/test_gpu_cloud.py tests/test_gpu_cpp_dropin.py tests/test_gpu_python_api.py -x -q -m gpu 2>&1 | tail -8
timeout 600 python tools/time_pose.py > gpurun_out/time_pose.log 2>&1; tail -5 gpurun_out/time_pose.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:cloud_tma_kernel -s 2 -c 1 -o gpurun_out/pose_k1 -f python tools/prof_pose.py > gpurun_out/ncu_pose.log 2>&1; tail -3 gpurun_out/ncu_pose.log
timeout 300 python tools/time_dewarp_frame.py 2>&1 | tail -15
