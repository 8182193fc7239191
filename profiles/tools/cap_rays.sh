# ncu capture of the ray kernel (C3) + AO-only / shadow-only timings; usage: bash profiles/tools/cap_rays.sh <tag>
TAG=${1:-x}; OUT=gpurun_out
timeout 300 python bench.py --config c3 --shadow-type 2 --no-e2e --no-cpu-baseline --steps 10 > $OUT/${TAG}_c3_aoonly.json 2> $OUT/${TAG}_err.txt
timeout 300 python bench.py --config c3 --no-e2e --no-cpu-baseline --steps 10 > $OUT/${TAG}_c3.json 2>> $OUT/${TAG}_err.txt
python profiles/bench_summary.py $OUT/${TAG}_c3_aoonly.json $OUT/${TAG}_c3.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_light_rays -s 2 -c 1 -f -o $OUT/${TAG}_k_light_rays \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_ncu.log 2>&1
