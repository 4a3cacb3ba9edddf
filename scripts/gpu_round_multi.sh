#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_mle2q_n2.json 2> gpurun_out/bench_mle2q_n2.err
cut -c1-600 gpurun_out/bench_mle2q_n2.json; tail -5 gpurun_out/bench_mle2q_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload pgdb3q --batch 296 --gpus 2 --steps 1 --warmup 3 > gpurun_out/bench_pgdb3q_n2.json 2> gpurun_out/bench_pgdb3q_n2.err
cut -c1-400 gpurun_out/bench_pgdb3q_n2.json; tail -5 gpurun_out/bench_pgdb3q_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
cut -c1-600 gpurun_out/bench_ref_n2.json; tail -5 gpurun_out/bench_ref_n2.err
