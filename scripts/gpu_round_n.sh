#!/bin/bash
set -x
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_n.log 2>&1
tail -4 gpurun_out/pytest_n.log
timeout 900 python bench.py --workload pgdb3q --batch 1024 --steps 2 --warmup 3 > gpurun_out/bench_pgdb3q_v9.json 2> gpurun_out/bench_pgdb3q_v9.err
cut -c1-300 gpurun_out/bench_pgdb3q_v9.json; tail -3 gpurun_out/bench_pgdb3q_v9.err
timeout 600 python bench.py --workload streaming > gpurun_out/bench_streaming_v3.json 2> gpurun_out/bench_streaming_v3.err
python -c "
import json
d=json.load(open('gpurun_out/bench_streaming_v3.json'))
for r in d['kernels']: print('%-70s %10.3f ms %8.1f GB/s %.3f' % (r['kernel'][:70], r['ms'], r['achieved_gbs'], r['frac_of_hbm_peak']))
"; tail -3 gpurun_out/bench_streaming_v3.err
cat > /tmp/race_small.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from forest_benchmarking_b200 import distance_measures as dm
from forest_benchmarking_b200.operator_tools import project_superoperators as pj
for d in (4, 16):
    g = torch.randn(3, d, d, dtype=torch.complex128, device="cuda")
    r = g @ g.conj().transpose(1, 2)
    dm.fidelity_batch(r, r.flip(0).contiguous())
pj.proj_choi_to_completely_positive_batch(torch.randn(2, 16, 16, dtype=torch.complex128, device="cuda"))
pj.proj_choi_to_completely_positive_batch(torch.randn(1, 64, 64, dtype=torch.complex128, device="cuda"))
pj.proj_choi_to_physical_batch(torch.randn(1, 16, 16, dtype=torch.complex128, device="cuda"))
torch.cuda.synchronize(); print("race workload done")
PY
timeout 420 compute-sanitizer --tool racecheck --print-limit 3 python /tmp/race_small.py > gpurun_out/sanitizer_racecheck_v2.log 2>&1
tail -4 gpurun_out/sanitizer_racecheck_v2.log | cut -c1-300
