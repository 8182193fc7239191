#!/bin/bash
# A/B of library variants under build/variants: usage AB_LIBS="a.so b.so" AB_CONFIGS="c3" AB_ENV="X=1" bash profiles/tools/ab_lib.sh <tag>
TAG=$1; OUT=gpurun_out; mkdir -p $OUT; : > $OUT/${TAG}_ab.jsonl
cp luz_b200/libluzrt.so /tmp/libluzrt_product.so
for lib in $AB_LIBS; do
  cp build/variants/$lib luz_b200/libluzrt.so
  for c in ${AB_CONFIGS:-c3}; do
    steps=20; [ $c = c4 ] && steps=5; [ $c = c5 ] && steps=5
    echo "## $lib $c" >> $OUT/${TAG}_ab.jsonl
    env ${AB_ENV:-X=0} timeout 300 python bench.py --config $c --no-e2e --no-cpu-baseline --no-parity --steps $steps --warmup 3 >> $OUT/${TAG}_ab.jsonl 2>> $OUT/${TAG}_ab.err
  done
done
cp /tmp/libluzrt_product.so luz_b200/libluzrt.so
python - <<PY
import json
for l in open("$OUT/${TAG}_ab.jsonl"):
    if l.startswith("##"): print(l.strip(), end="  ")
    elif l.startswith("{"):
        d = json.loads(l); k = d["kernels_ms"]; print("ms/step %.3f light %.3f rays %.3f shade %.3f taa %.3f" % (d["ms_per_step"], k["light"], k["light_rays"], k["light_shade"], k["taa"]))
PY
