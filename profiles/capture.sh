#!/bin/bash
# One GPU call: parity tests, bench lines, ncu launch list and full captures of the hot kernels.
# usage (from the repo root): gpurun --timeout 2400 -- 'bash profiles/capture.sh <tag>'
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
: > $OUT/${TAG}_bench_other.json
for c in c1 c2 c4 c5; do steps=10; [ $c = c5 ] && steps=5; timeout 400 python bench.py --config $c --no-cpu-baseline --steps $steps >> $OUT/${TAG}_bench_other.json 2>> $OUT/${TAG}_bench.err; done
timeout 400 python bench.py --config c3 --variant unique --no-cpu-baseline --no-e2e >> $OUT/${TAG}_bench_other.json 2>> $OUT/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches_bench.log 2>&1
for k in k_light_rays_split k_light_shade k_taa k_shadow_hints; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $OUT/${TAG}_$k \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_ncu_$k.log 2>&1
done
tail -3 $OUT/${TAG}_pytest.log; cat $OUT/${TAG}_bench.json
