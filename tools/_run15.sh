set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cloud.py tests/test_gpu_dewarp.py tests/test_gpu_edge_cases.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python tools/time_pose.py > gpurun_out/time_pose.log 2>&1; tail -60 gpurun_out/time_pose.log | head -90
for f in 64 128; do
timeout 600 python bench.py --only k2 --kernel-only --k2-frames $f --steps 20 --warmup 5 > gpurun_out/k2_f$f.json 2> gpurun_out/k2_f$f.err
tail -c 600 gpurun_out/k2_f$f.json
done
