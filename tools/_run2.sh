set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_pipe -s 6 -c 1 -o gpurun_out/k2_pipe_r02 -f python bench.py --only k2 --kernel-only --steps 3 --warmup 3 > gpurun_out/ncu_k2.log 2>&1
tail -3 gpurun_out/ncu_k2.log
for w in 24 18; do
OB_DECODE_PIPE_WARPS=$w timeout 300 python bench.py --only k2 --kernel-only --steps 20 --warmup 5 2>&1 | tail -1
done
