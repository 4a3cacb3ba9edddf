#!/bin/bash
# round 2, call V: ncu --set full + source view of the compact fidelity_tri_kernel<16>
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2v_build.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fidelity_tri_kernel -s 2 -c 1 -o gpurun_out/r2v_prof_fid -f python scripts/prof_fid.py 4 > gpurun_out/r2v_ncu.log 2>&1
tail -2 gpurun_out/r2v_ncu.log
python scripts/summarize_ncu.py full gpurun_out/r2v_prof_fid.ncu-rep gpurun_out/r2v_ncu_fid.md fidelity_tri
python scripts/ncu_lines.py gpurun_out/r2v_prof_fid.ncu-rep 60 > gpurun_out/r2v_ncu_fid_lines.txt 2>&1
head -5 gpurun_out/r2v_ncu_fid_lines.txt
