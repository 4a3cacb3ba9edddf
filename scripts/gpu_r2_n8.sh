#!/bin/bash
# round 2: the default bench line under torchrun on all GPUs of the box (what the driver's scaling run does)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2n8_build.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --no-cpu-baseline > gpurun_out/r2n8_bench.json 2> gpurun_out/r2n8_bench.err; echo "bench N=$N rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2n8_bench.json") if l.startswith("{")][0])
print("mle2q", round(d["value"]), "n_gpus", d["n_gpus"], "pgdb3q", round(d["pgdb3q"]["value"], 1), "distances", round(d["distances"]["value"]), "mle3q", round(d["mle3q"]["value"]))
PY
tail -3 gpurun_out/r2n8_bench.err
