#!/bin/bash
# A/B of prebuilt libluzrt.so variants on one box: usage  bash profiles/ab.sh <tag> <variant.so>...   (last one stays installed)
# Each variant is copied over luz_b200/libluzrt.so and benched on C3 / C4 / C2 (device-resident timing only).
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/${TAG}_ab.jsonl
for so in "$@"; do
  cp "$so" luz_b200/libluzrt.so
  for env in ${AB_ENVS:-"X=0"}; do
    for c in ${AB_CONFIGS:-c3 c4 c2}; do
      steps=20; [ $c = c4 ] && steps=5
      echo "## $so $env $c" >> $OUT/${TAG}_ab.jsonl
      env $env timeout 300 python bench.py --config $c $AB_EXTRA --no-e2e --no-cpu-baseline --steps $steps --warmup 3 >> $OUT/${TAG}_ab.jsonl 2>> $OUT/${TAG}_ab.err
    done
  done
done
python - <<PY
import json
for l in open("$OUT/${TAG}_ab.jsonl"):
    if l.startswith("##"): print(l.strip(), end="  ")
    elif l.startswith("{"):
        d = json.loads(l); print("ms/step %.3f light %.3f taa %.3f" % (d["ms_per_step"], d["kernels_ms"]["light"], d["kernels_ms"]["taa"]))
PY
