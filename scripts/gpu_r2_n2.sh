#!/bin/bash
# round 2: the default bench line under torchrun on 2 GPUs (MLE weak scaling, PGDB / distances strong scaling + all-gather)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2n2_build.log 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r2n2_bench.json 2> gpurun_out/r2n2_bench.err; echo "bench N=2 rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2n2_bench.json") if l.startswith("{")][0])
print("mle2q", round(d["value"]), "n_gpus", d["n_gpus"], "pgdb3q", round(d["pgdb3q"]["value"], 1), "distances", round(d["distances"]["value"]))
PY
timeout 600 python -m pytest tests -m gpu -x -q -k "allgather or shard or multi" > gpurun_out/r2n2_pytest.log 2>&1; tail -2 gpurun_out/r2n2_pytest.log
