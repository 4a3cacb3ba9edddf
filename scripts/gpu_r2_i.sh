#!/bin/bash
# round 2, call I: register-resident n = 3 PTM kernel: parity (convert + process tests use it), timing vs the shared-memory stages
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2i_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_convert.py -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2i_pytest.log
timeout 300 python scripts/prof_ptm_dense.py > gpurun_out/r2i_ptm.json 2> gpurun_out/r2i_ptm.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2i_ptm.json"))
for r in d["rows"]:
    print(r["n"], r["variant"], r["direction"], round(r["ms"], 3), "ms", round(r["frac_of_hbm_peak"], 3))
PY
