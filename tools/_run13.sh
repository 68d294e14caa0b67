set -x
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitize_cases.py > gpurun_out/r02_sanitize_$tool.txt 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/r02_sanitize_$tool.txt
done
