set -x
mkdir -p gpurun_out
timeout 600 python tools/k2_parts.py 2>&1 | tail -9
