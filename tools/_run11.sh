set -x
mkdir -p gpurun_out
timeout 900 python tools/sweep_k1_single.py 2>&1 | tail -30
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.json
