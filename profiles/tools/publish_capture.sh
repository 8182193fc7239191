#!/bin/bash
# Copies a capture_r2.sh run (gpurun_out/<tag>_*) into profiles/ under r2_*_<ver> names and rewrites traffic.json.
# usage: bash profiles/tools/publish_capture.sh <tag> <ver>      e.g.  r4 v3
TAG=$1; VER=$2; G=gpurun_out; P=profiles
for k in k_shadow_hints k_shadow_rays_temporal k_ao_rays_compact k_light_shade k_taa; do
  python $P/summarize.py $G/${TAG}_$k.ncu-rep "r2 $VER $k (C3 4K, 1xB200)" > $P/r2_${k}_$VER.md
done
python $P/tools/make_traffic.py c3:k_shadow_hints:$G/${TAG}_k_shadow_hints.ncu-rep c3:k_shadow_rays_temporal:$G/${TAG}_k_shadow_rays_temporal.ncu-rep \
    c3:k_ao_rays_compact:$G/${TAG}_k_ao_rays_compact.ncu-rep c3:k_light_shade:$G/${TAG}_k_light_shade.ncu-rep c3:k_taa:$G/${TAG}_k_taa.ncu-rep
cp $G/${TAG}_bench.json $P/r2_bench_c3_$VER.json
cp $G/${TAG}_bench_other.json $P/r2_bench_other_$VER.json
cp $G/${TAG}_bench_ref.json $P/r2_bench_ref_$VER.json
cp $G/${TAG}_launches.csv $P/r2_launches_c3_$VER.csv
cp $G/${TAG}_pytest.log $P/r2_gpu_pytest_$VER.log
cp $G/${TAG}_pytest_exact_raygen.log $P/r2_gpu_pytest_exact_raygen_$VER.log
cp $G/${TAG}_sanitizer_memcheck.log $P/r2_sanitizer_memcheck.log
cp $G/${TAG}_sanitizer_racecheck.log $P/r2_sanitizer_racecheck.log
cat $G/${TAG}_smi.txt
