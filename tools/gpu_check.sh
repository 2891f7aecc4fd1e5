#!/usr/bin/env bash
# GPU box helper: parity suite + short bench lines.  usage: tools/gpu_check.sh <tag> [pytest -k expr] [workloads...]
tag=${1:-run}; kexpr=${2:-}; shift; shift || true
wls=${@:-config2 config3}
mkdir -p gpurun_out
if [ "$kexpr" != "skip" ]; then
  if [ -n "$kexpr" ]; then timeout 1200 python -m pytest tests -m gpu -x -q -k "$kexpr" 2>&1 | tail -12; else timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12; fi
fi
for w in $wls; do
  timeout 400 python bench.py --workload $w --steps 3 --no-cpu-baseline --no-dropin --no-secondary > gpurun_out/${tag}_$w.json 2> gpurun_out/${tag}_$w.err; echo "$w rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_$w.json"))
    print("$w", round(d["value"]), "rays/s", round(d["ms_per_step"], 1), "ms/step  e2e", round(d["e2e"]["value"]))
    print("   ", {k: round(v, 1) for k, v in d["roofline"]["kernels_ms_per_step"].items() if v > 1})
except Exception as e:
    print("$w: no json", e); print(open("gpurun_out/${tag}_$w.err").read()[-2000:])
PY
done
