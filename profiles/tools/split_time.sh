#!/bin/bash
# Time split of the C3 ray kernel: shadow rays only, AO rays only, both (device-resident timing).
OUT=gpurun_out; TAG=${1:-split}
: > $OUT/${TAG}.jsonl
for extra in "" "--ao-samples 0" "--light-samples 0"; do
  echo "## c3 $extra" >> $OUT/${TAG}.jsonl
  timeout 300 python bench.py --config c3 $extra --no-e2e --no-cpu-baseline --no-parity --steps 20 --warmup 3 >> $OUT/${TAG}.jsonl 2>> $OUT/${TAG}.err
done
python - <<PY
import json
for l in open("$OUT/${TAG}.jsonl"):
    if l.startswith("##"): print(l.strip(), end="  ")
    elif l.startswith("{"):
        d = json.loads(l); k = d["kernels_ms"]; print("ms/step %.3f light %.3f rays %.3f shade %.3f taa %.3f" % (d["ms_per_step"], k["light"], k["light_rays"], k["light_shade"], k["taa"]))
PY
