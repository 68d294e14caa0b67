set -x
mkdir -p gpurun_out
NCCL_DEBUG=INFO timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.out 2> gpurun_out/bench_n2.err
echo rc=$?; tail -c 400 gpurun_out/bench_n2.err; grep -c "NCCL INFO" gpurun_out/bench_n2.out; tail -n 1 gpurun_out/bench_n2.out | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 2 > gpurun_out/bench_ref_n2.out 2> gpurun_out/bench_ref_n2.err; tail -n 1 gpurun_out/bench_ref_n2.out | cut -c1-300
