set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_pcap_source.py tests/test_gpu_batcher.py -x -q -m gpu 2>&1 | tail -4
for tool in racecheck memcheck; do
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitize_cases.py > gpurun_out/r02_sanitize_$tool.txt 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/r02_sanitize_$tool.txt
done
timeout 600 python bench.py --only k2 --kernel-only --steps 20 --warmup 5 > gpurun_out/k2_only.json 2> gpurun_out/k2_only.err
tail -c 1500 gpurun_out/k2_only.json
