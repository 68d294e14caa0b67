set -x
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|Thread"
for envs in "X=1" "OMP_WAIT_POLICY=passive" "OMP_WAIT_POLICY=active" "OMP_PROC_BIND=true OMP_PLACES=cores" "OMP_NUM_THREADS=64" "GOMP_SPINCOUNT=0"; do
  echo "== $envs"
  env $envs timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(round(d['value']), d['cpu_baseline']['mode'], {k:round(v) for k,v in d['cpu_baseline']['modes'].items()})"
done
