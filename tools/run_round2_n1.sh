# Round-2 validation on one GPU (run under gpurun): GPU test-suite, ncu traffic records, bench + CPU arm, launch list, sanitizers, K2 part timing
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python tools/capture_traffic.py 2>&1 | tail -4
cp gpurun_out/k1_traffic.json gpurun_out/k2_traffic.json gpurun_out/k2_streams8_traffic.json profiles/
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err; tail -c 200 gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_n1.json 2>gpurun_out/bench_ref_n1.err; tail -c 200 gpurun_out/bench_ref_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sweep > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-200
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitize_cases.py > gpurun_out/r02_sanitize_$tool.txt 2>&1
  echo "$tool rc=$?"; tail -2 gpurun_out/r02_sanitize_$tool.txt
done
timeout 600 python tools/k2_parts.py 2>&1 | tail -9
timeout 600 ncu --set full --import-source on --clock-control none -k regex:decode_pipe -s 4 -c 1 -o gpurun_out/k2_pipe_head -f python bench.py --only k2 --kernel-only --steps 3 --warmup 3 > gpurun_out/ncu_k2_head.log 2>&1; tail -1 gpurun_out/ncu_k2_head.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:cloud_tma -s 4 -c 1 -o gpurun_out/k1_head -f python bench.py --only k1 --kernel-only --steps 3 --warmup 3 > gpurun_out/ncu_k1_head.log 2>&1; tail -1 gpurun_out/ncu_k1_head.log
