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
timeout 400 python bench.py --config c3 --variant unique --no-cpu-baseline >> $OUT/${TAG}_bench_other.json 2>> $OUT/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
bash profiles/f4_bench.sh > $OUT/${TAG}_f4_summary.txt 2>&1; cp $OUT/r1f_f4.jsonl $OUT/${TAG}_bench_f4.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches_bench.log 2>&1
for k in k_light_rays k_light_shade k_taa; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $OUT/${TAG}_$k \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_ncu_$k.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shadow_map -s 2 -c 1 -f -o $OUT/${TAG}_k_shadow_map \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --shadow-type 2 > $OUT/${TAG}_ncu_k_shadow_map.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_volumetric_screen -s 2 -c 1 -f -o $OUT/${TAG}_k_volumetric_screen \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --volumetric 1 > $OUT/${TAG}_ncu_k_volumetric_screen.log 2>&1
tail -3 $OUT/${TAG}_pytest.log; cat $OUT/${TAG}_bench.json
