# 8-GPU validation of the bench contract (run under gpurun --gpus 8); lines -> gpurun_out/
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/n${N}_k1.err | tail -1 > gpurun_out/bench_k1_n$N.json
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --workload k2 --streams-per-gpu 8 2>gpurun_out/n${N}_k2.err | tail -1 > gpurun_out/bench_k2_n$N.json
timeout 400 $TR bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>gpurun_out/n${N}_ref.err | tail -1 > gpurun_out/bench_ref_n$N.json
for f in k1 k2 ref; do head -c 600 gpurun_out/bench_${f}_n$N.json; echo; done
