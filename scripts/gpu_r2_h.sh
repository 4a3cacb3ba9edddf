#!/bin/bash
# round 2, call H (2 GPUs): the driver's launch line for N = 2 (strong-scaling PGDB + distances inside the default bench)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2h_build.log 2>&1
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err; echo "bench n2 rc=$?"
tail -c 300 gpurun_out/r2h_bench_n2.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2h_bench_n2.json"))
print("mle", d["value"], d["n_gpus"], "pgdb3q", d["pgdb3q"]["value"], d["pgdb3q"]["kernel_ms_per_rank"], "dist", d["distances"]["value"])
PY
timeout 300 python -m pytest tests/test_gpu_state.py -m gpu -x -q -k "another_device" > gpurun_out/r2h_pytest_2gpu.log 2>&1; tail -2 gpurun_out/r2h_pytest_2gpu.log
