set -x
mkdir -p gpurun_out
K2_DYN="3,0" K2_LANE_ARRIVE="1,0" K2_VARIANTS="none,fields_only,all" timeout 900 python tools/k2_parts.py 2>&1 | tail -14
