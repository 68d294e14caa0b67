set -x
mkdir -p gpurun_out
NCCL_DEBUG=INFO timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_n4.out 2> gpurun_out/bench_n4.err
echo rc=$?; grep -c "NCCL INFO" gpurun_out/bench_n4.out; tail -n 1 gpurun_out/bench_n4.out | cut -c1-400
nvidia-smi topo -m > gpurun_out/topo_n4.txt 2>&1
