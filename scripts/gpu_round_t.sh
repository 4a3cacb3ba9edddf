#!/bin/bash
set -x
mkdir -p gpurun_out
for w in moments fidelity pgdb2; do
  case $w in moments) k=moments_swar_kernel;; fidelity) k=fidelity_kernel;; *) k=pgdb_kernel;; esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/prof_$w -f python scripts/prof_next.py $w > gpurun_out/ncu_$w.log 2>&1
  tail -2 gpurun_out/ncu_$w.log
done
ls -la gpurun_out/*.ncu-rep
