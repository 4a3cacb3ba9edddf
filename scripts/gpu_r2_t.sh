#!/bin/bash
# round 2, call T: does block-level lockstep cut instruction-cache traffic?  ncu metrics of two ubench variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SRC="scripts/ubench_fid.cu forest_benchmarking_b200/csrc/qt_distance.cu forest_benchmarking_b200/csrc/qt_api.cu"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DFID_TRI_WPB_N=4 -o /tmp/fa.bin $SRC > /dev/null 2>&1 &
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DFID_TRI_WPB_N=12 -DFID_TRI_LB_T=384 -DFID_TRI_LB_B=1 -DFID_TRI_SYNC=1 -o /tmp/fb.bin $SRC > /dev/null 2>&1 &
wait
M=gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__issue_active.avg.per_cycle_active,smsp__inst_executed.sum,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warp_latency_per_inst_issued.ratio,sm__icc_requests.sum,sm__icc_requests_lookup_miss.sum,sm__icc_requests_lookup_hit.sum
for b in fa fb; do
  timeout 300 ncu --metrics $M --clock-control none -k regex:fidelity_tri_kernel -s 2 -c 1 --csv --log-file gpurun_out/r2t_$b.csv /tmp/$b.bin 4 262144 > /dev/null 2>&1
  echo "== $b"; python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2t_$b.csv")) if len(r) > 5]
for r in rows[1:]: print(r[-3], r[-1])
PY
done
ncu --query-metrics 2>/dev/null | grep -i "icc\|inst_fetch\|l0" | head -20
