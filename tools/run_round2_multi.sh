# N-GPU bench lines of round 2 (run under `gpurun --gpus N -- bash tools/run_round2_multi.sh N`); NCCL INFO lines kept
set -x
N=${1:-2}
mkdir -p gpurun_out
NCCL_DEBUG=INFO timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.out 2> gpurun_out/bench_n$N.err
echo rc=$?; grep -c "NCCL INFO" gpurun_out/bench_n$N.out; tail -n 1 gpurun_out/bench_n$N.out | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 5 --warmup 2 > gpurun_out/bench_ref_n$N.out 2> gpurun_out/bench_ref_n$N.err; tail -n 1 gpurun_out/bench_ref_n$N.out | cut -c1-300
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
