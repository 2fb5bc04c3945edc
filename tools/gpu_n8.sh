#!/bin/bash
# Multi-GPU evidence on one box (run under `gpurun --gpus N`): bench.py at N ranks (weak scaling by
# streams) and the NCCL scatter/gather of BASELINE config 5 at full size (1 048 576 stereo frames).
N=${1:-8}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -c 1500 gpurun_out/bench_n$N.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 3 --no-e2e --workload config5 > gpurun_out/bench_config5_n$N.json 2>> gpurun_out/bench_n$N.err; tail -c 600 gpurun_out/bench_config5_n$N.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/scatter_bench.py --streams 4096 --frames 256 --steps 5 > gpurun_out/scatter_n$N.json 2>> gpurun_out/bench_n$N.err; tail -c 800 gpurun_out/scatter_n$N.json
tail -3 gpurun_out/bench_n$N.err
