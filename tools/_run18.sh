set -x
mkdir -p gpurun_out
timeout 600 python tools/k2_parts.py 2>&1 | tail -8
timeout 600 python tools/time_pose.py > gpurun_out/time_pose.log 2>&1; python - <<'PY'
import json
d=json.load(open('gpurun_out/time_pose.json'))
for r in d['sweep_best']: print(r)
print({k:v for k,v in d.items() if k.endswith('_ms')})
PY
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitize_cases.py > gpurun_out/r02_sanitize_$tool.txt 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/r02_sanitize_$tool.txt
done
