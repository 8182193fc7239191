#!/bin/bash
# A/B of environment switches on the installed libluzrt.so: usage  AB_ENVS="A=1 B=2" AB_CONFIGS="c3 c4" bash profiles/tools/ab_env.sh <tag>
TAG=$1
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/${TAG}_ab.jsonl
for env in ${AB_ENVS:-"X=0"}; do
  for c in ${AB_CONFIGS:-c3 c4 c2}; do
    steps=20; [ $c = c4 ] && steps=5; [ $c = c5 ] && steps=5
    echo "## $env $c" >> $OUT/${TAG}_ab.jsonl
    env $env timeout 300 python bench.py --config $c $AB_EXTRA --no-e2e --no-cpu-baseline --no-parity --steps $steps --warmup 3 >> $OUT/${TAG}_ab.jsonl 2>> $OUT/${TAG}_ab.err
  done
done
python - <<PY
import json
for l in open("$OUT/${TAG}_ab.jsonl"):
    if l.startswith("##"): print(l.strip(), end="  ")
    elif l.startswith("{"):
        d = json.loads(l); k = d["kernels_ms"]; print("ms/step %.3f light %.3f rays %.3f shade %.3f taa %.3f" % (d["ms_per_step"], k["light"], k["light_rays"], k["light_shade"], k["taa"]))
PY
