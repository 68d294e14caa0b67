#!/bin/bash
# Tunable sweep of the K1 kernel (run under gpurun). Each config is its own process.
out=gpurun_out/sweep_k1.jsonl; : > $out
for tw in 256 512 1024; do for st in 3 4 6; do for cta in 2 3 4 6; do for th in 64 128 256; do
  smem=$(( (tw*32 + 127) * st ))
  OB_CLOUD_TW=$tw OB_CLOUD_STAGES=$st OB_CLOUD_CTAS_PER_SM=$cta OB_CLOUD_THREADS=$th \
    timeout 120 python bench.py --kernel-only --steps 10 --warmup 3 2>/dev/null | tail -1 >> $out
done; done; done; done
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/sweep_k1.jsonl') if l.strip().startswith('{')]
rows.sort(key=lambda r:-r['gbps'])
for r in rows[:15]: print(round(r['gbps']), round(r['frac'],3), r['env'])
print('worst', round(rows[-1]['gbps']), rows[-1]['env'])
PY
