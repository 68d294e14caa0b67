# ncu evidence for profiles/: launch lists of the bench commands + one full capture per kernel
# (run under gpurun; numbers printed by a bench under ncu are never bench values)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/k1_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/k1_bench_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/k2_bench_launches.csv python bench.py --workload k2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/k2_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cloud_tma_kernel -s 2 -c 1 -o gpurun_out/k1_prof_v3 python bench.py --no-cpu-baseline --kernel-only --steps 3 --warmup 2 > gpurun_out/k1_prof_v3.log 2>&1
tail -2 gpurun_out/k1_prof_v3.log
wc -l gpurun_out/k1_bench_launches.csv gpurun_out/k2_bench_launches.csv
