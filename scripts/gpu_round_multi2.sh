#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_mle2q_n2.json 2> gpurun_out/bench_mle2q_n2.err
cut -c1-700 gpurun_out/bench_mle2q_n2.json; tail -3 gpurun_out/bench_mle2q_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
cut -c1-300 gpurun_out/bench_ref_n2.json; tail -3 gpurun_out/bench_ref_n2.err
