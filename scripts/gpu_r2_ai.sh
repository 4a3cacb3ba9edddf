#!/bin/bash
# round 2, call AI: TNI n = 3 correction pass at 2 / 3 / 4 resident blocks per SM (register cap)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for mb in 2 3 4; do
  sed -i "s/#define QT_TNI3_MINB [0-9]/#define QT_TNI3_MINB $mb/" forest_benchmarking_b200/csrc/qt_project.cu
  python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2ai_build_$mb.log 2>&1
  timeout 600 python bench.py --workload streaming --no-cpu-baseline > gpurun_out/r2ai_streaming_$mb.json 2> gpurun_out/r2ai_streaming_$mb.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r2ai_streaming_$mb.json"))
for r in d["kernels"]:
    if "tni_correction_kernel<3>" in r["kernel"]: print("MINB=$mb  %.2f %8.3f ms" % (r["frac_of_hbm_peak"], r["ms"]))
PY
done
