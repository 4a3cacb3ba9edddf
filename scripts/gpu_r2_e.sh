#!/bin/bash
# round 2, call E: role-split Jacobi ablations; dense FP64-MMA PTM variant: parity, timing, tensor-pipe evidence
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 scripts/ubench_jacobi.bin > gpurun_out/r2e_ubench_jacobi.txt 2>&1; echo "ubench rc=$?"
grep -i "split\|plain" gpurun_out/r2e_ubench_jacobi.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2e_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_convert.py -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2e_pytest.log
timeout 300 python scripts/prof_ptm_dense.py > gpurun_out/r2e_ptm_dense.json 2> gpurun_out/r2e_ptm_dense.err; cat gpurun_out/r2e_ptm_dense.json | head -c 3000
ncu --query-metrics 2>/dev/null | grep -i "dmma\|pipe_tensor" | head -20 > gpurun_out/r2e_ncu_metric_names.txt
timeout 600 ncu --set full --clock-control none -k regex:"pl_dense_dmma_kernel|pl_pass_a_kernel" -c 12 -o gpurun_out/r2e_prof_ptm -f python scripts/prof_ptm_dense.py 1 > gpurun_out/r2e_ncu.log 2>&1
tail -2 gpurun_out/r2e_ncu.log
python scripts/summarize_ncu.py full gpurun_out/r2e_prof_ptm.ncu-rep gpurun_out/r2e_ncu_ptm.md
