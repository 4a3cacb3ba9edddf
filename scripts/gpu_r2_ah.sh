#!/bin/bash
# round 2, call AH: d = 32 tridiagonal fidelity kernel -- state tests + accuracy / timing sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2ah_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_state.py -m gpu -x -q > gpurun_out/r2ah_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2ah_pytest.log
timeout 600 python scripts/fid_accuracy.py > gpurun_out/r2ah_fid_accuracy.txt 2>&1; grep "n=5\|B=" gpurun_out/r2ah_fid_accuracy.txt
