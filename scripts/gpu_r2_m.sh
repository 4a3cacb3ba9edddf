#!/bin/bash
# round 2, call M: Choi projections at n = 4, 5; full GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2m_build.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_project.py -m gpu -x -q -k large > gpurun_out/r2m_pytest_large.log 2>&1; echo "pytest large rc=$?"; tail -15 gpurun_out/r2m_pytest_large.log
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_project.py::test_large_choi_projections > gpurun_out/r2m_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -4 gpurun_out/r2m_pytest_all.log
