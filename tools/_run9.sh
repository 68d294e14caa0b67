set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python tools/capture_traffic.py 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err; tail -c 400 gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_n1.json 2>gpurun_out/bench_ref_n1.err; tail -c 300 gpurun_out/bench_ref_n1.json
