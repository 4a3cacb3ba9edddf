#!/bin/bash
# round 2, call F: robustness of the early-stop threshold (trip counts vs tight run); ncu of the dense PTM variant at n = 3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2f_build.log 2>&1
timeout 600 python scripts/exp_pgdb_tol.py 296 > gpurun_out/r2f_exp_pgdb_tol.txt 2>&1; tail -9 gpurun_out/r2f_exp_pgdb_tol.txt | cut -c1-400
timeout 600 ncu --set full --clock-control none -k regex:"pl_dense_dmma_kernel" -s 6 -c 2 -o gpurun_out/r2f_prof_ptm3 -f python scripts/prof_ptm_dense.py 1 > gpurun_out/r2f_ncu.log 2>&1
python scripts/summarize_ncu.py full gpurun_out/r2f_prof_ptm3.ncu-rep gpurun_out/r2f_ncu_ptm3_dense.md
grep "^## \|tensor\|gpu__time" gpurun_out/r2f_ncu_ptm3_dense.md
