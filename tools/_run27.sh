set -x
./tools/micro/stg_issue
