#!/bin/bash
# round 2, call U: compact (rolled, shared-memory) fidelity_tri_kernel -- distance tests, accuracy sweep, block-shape variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2u_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_state.py tests/test_gpu_convert.py -m gpu -x -q -k "distance or fidelity or empty or purity" > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2u_pytest.log
timeout 600 python scripts/fid_accuracy.py > gpurun_out/r2u_fid_accuracy.txt 2>&1; grep "n=4\|n=3 \|B=" gpurun_out/r2u_fid_accuracy.txt
bash scripts/gpu_r2_s.sh
