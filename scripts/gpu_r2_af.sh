#!/bin/bash
# round 2, call AF: wall-clock of the two driver arms
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2af_build.log 2>&1
( time python bench.py > gpurun_out/r2af_bench.json 2> gpurun_out/r2af_bench.err ) 2> gpurun_out/r2af_time_ours.txt; cat gpurun_out/r2af_time_ours.txt
( time python bench.py --impl reference > gpurun_out/r2af_bench_ref.json 2> gpurun_out/r2af_bench_ref.err ) 2> gpurun_out/r2af_time_ref.txt; cat gpurun_out/r2af_time_ref.txt
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2af_bench.json") if l.startswith("{")][0])
print(d["value"], d["pgdb3q"]["value"], d["distances"]["value"], d["distances"]["roofline"]["traffic"], d["mle3q"]["value"])
PY
