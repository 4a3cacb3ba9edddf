#!/bin/bash
# DRAM traffic of every HBM-bound kernel (ncu, 3 metrics only): compare with the algorithmic bytes of bench_kernels.py
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/traffic_streaming.csv python bench.py --workload streaming > gpurun_out/traffic_streaming.json 2> gpurun_out/traffic_streaming.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/traffic_convert.csv python bench.py --workload convert > gpurun_out/traffic_convert.json 2> gpurun_out/traffic_convert.err
ls -la gpurun_out/traffic_*
