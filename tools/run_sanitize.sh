# compute-sanitizer passes over tools/sanitize_cases.py (run under gpurun); logs -> gpurun_out/
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_cases.py > gpurun_out/sanitize_$tool.txt 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/sanitize_$tool.txt
done
