# full GPU regression + every bench line + the sweeps that feed profiles/ (run under gpurun)
set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_k1.json
timeout 300 python bench.py --workload k2 2>/dev/null | tail -1 > gpurun_out/bench_k2.json
timeout 300 python bench.py --workload k2 --streams-per-gpu 8 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_k2_ms.json
timeout 300 python tools/sweep_shapes.py 2>&1 | tail -30
timeout 200 python tools/sweep_batch.py 2>&1 | tail -26
