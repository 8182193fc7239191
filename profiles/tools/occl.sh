OUT=gpurun_out
for c in "c3" "c3 --shadow-type 2" "c4" "c2"; do
  timeout 300 python bench.py --config $c --no-e2e --no-cpu-baseline --steps 5 --warmup 3 > $OUT/occl_tmp.json 2>>$OUT/occl.err
  echo "## $c"; python profiles/bench_summary.py $OUT/occl_tmp.json
done
