#!/bin/bash
# round 2, call S: fidelity_tri_kernel block-shape variants (scripts/ubench_fid_variants.txt; the barrier / staging / rolled
# variants of profiles/r02_ubench_fidelity_tri.txt are in the git history of qt_distance.cu), n = 4, 2^18 pairs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SRC="scripts/ubench_fid.cu forest_benchmarking_b200/csrc/qt_distance.cu forest_benchmarking_b200/csrc/qt_api.cu"
i=0
while read -r flags; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 $flags -o /tmp/ubench_fid_$i.bin $SRC > /dev/null 2>&1 &
  i=$((i+1))
done < scripts/ubench_fid_variants.txt
wait
i=0
while read -r flags; do
  echo "[$flags]"; timeout 120 /tmp/ubench_fid_$i.bin ${UB_N:-4} ${UB_B:-262144}
  i=$((i+1))
done < scripts/ubench_fid_variants.txt | tee gpurun_out/r2s_ubench_fid.txt
