#!/bin/bash
# round 2, final evidence run: what the driver does at round end (GPU tests, smoke, both bench arms), then the ncu evidence
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r2final
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${T}_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "ref arm rc=$?"
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/${T}_bench.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/${T}_bench.json") if l.startswith("{")][0])
print("mle2q", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), "parity", d.get("parity", {}).get("max_rel_frobenius_err"), d.get("parity", {}).get("iteration_count_mismatches"))
p = d["pgdb3q"]; print("pgdb3q", round(p["value"], 1), "e2e", round(p["e2e"]["value"], 1), "frac", round(p["roofline"]["frac"], 3), "parity", p.get("parity"))
print("distances", round(d["distances"]["value"]), "mle3q", round(d["mle3q"]["value"]))
PY
# ncu evidence: launch list of the bench command, then --set full of the two dominant kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_launch.log 2>&1
python scripts/summarize_ncu.py launches gpurun_out/${T}_launches.csv gpurun_out/${T}_launches.md; head -12 gpurun_out/${T}_launches.md
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mle_quad_kernel -s 1 -c 1 -o gpurun_out/${T}_prof_mle_quad -f python bench.py --workload mle2q --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_mle.log 2>&1
python scripts/summarize_ncu.py full gpurun_out/${T}_prof_mle_quad.ncu-rep gpurun_out/${T}_ncu_mle_quad.md mle_quad
timeout 900 ncu --set full --clock-control none -k regex:pgdb_kernel -s 1 -c 1 -o gpurun_out/${T}_prof_pgdb3 -f python bench.py --workload pgdb3q --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_pgdb.log 2>&1
python scripts/summarize_ncu.py full gpurun_out/${T}_prof_pgdb3.ncu-rep gpurun_out/${T}_ncu_pgdb3.md pgdb_kernel
timeout 600 ncu --set full --clock-control none -k regex:"fidelity_tri_kernel" -s 3 -c 1 -o gpurun_out/${T}_prof_misc -f python bench.py --workload distances --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_misc.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:mle_warp_kernel -s 1 -c 1 -o gpurun_out/${T}_prof_mle3q -f python bench.py --workload mle3q --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_mle3q.log 2>&1
python scripts/summarize_ncu.py full gpurun_out/${T}_prof_mle3q.ncu-rep gpurun_out/${T}_ncu_mle_warp3.md mle_warp
python scripts/summarize_ncu.py full gpurun_out/${T}_prof_misc.ncu-rep gpurun_out/${T}_ncu_fidelity_tri.md fidelity_tri
python scripts/make_traffic_json.py gpurun_out/${T}_traffic.json "mle_quad_kernel=gpurun_out/${T}_prof_mle_quad.ncu-rep:mle_quad" "pgdb_kernel<3>=gpurun_out/${T}_prof_pgdb3.ncu-rep:pgdb_kernel" "fidelity_tri_kernel<16>=gpurun_out/${T}_prof_misc.ncu-rep:fidelity_tri" "mle_warp_kernel<3>=gpurun_out/${T}_prof_mle3q.ncu-rep:mle_warp" | head -30
for w in streaming convert next; do timeout 900 python bench.py --workload $w > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; echo "$w rc=$?"; done
