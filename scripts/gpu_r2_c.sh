#!/bin/bash
# round 2, call C: ncu --set full + source view of pgdb_kernel<3> (second-generation Dykstra), one wave
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2c_build.log 2>&1
python scripts/prof_pgdb3.py 148
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pgdb_kernel -s 1 -c 1 -o gpurun_out/r2c_prof_pgdb3 -f python scripts/prof_pgdb3.py 148 > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
python scripts/summarize_ncu.py full gpurun_out/r2c_prof_pgdb3.ncu-rep gpurun_out/r2c_ncu_pgdb3.md pgdb_kernel
python scripts/ncu_lines.py gpurun_out/r2c_prof_pgdb3.ncu-rep 60 > gpurun_out/r2c_ncu_pgdb3_lines.txt 2>&1
head -5 gpurun_out/r2c_ncu_pgdb3_lines.txt
ls -la gpurun_out/r2c_prof_pgdb3.ncu-rep
