#!/bin/bash
# Round-2 capture on one B200: parity suites, bench lines of every config, reference arm, ncu launch list and full
# captures of the five kernels of a C3 frame.   usage: gpurun --timeout 3000 -- 'bash profiles/capture_r2.sh <tag>'
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1; nproc >> $OUT/${TAG}_smi.txt
timeout 1200 python -m pytest tests -m gpu -q -s > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
if [ -f build/variants/libluzrt_exact_raygen.so ]; then  # ray generation with IEEE sqrt / division / sincosf (LUZ_FAST_RAYGEN=0)
  cp luz_b200/libluzrt.so /tmp/libluzrt_product.so; cp build/variants/libluzrt_exact_raygen.so luz_b200/libluzrt.so
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_parity.py -m gpu -q > $OUT/${TAG}_pytest_exact_raygen.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_exact_raygen.log
  cp /tmp/libluzrt_product.so luz_b200/libluzrt.so
fi
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
: > $OUT/${TAG}_bench_other.json
for c in c1 c2 c4 c5; do steps=10; [ $c = c5 ] && steps=5; timeout 500 python bench.py --config $c --no-cpu-baseline --steps $steps >> $OUT/${TAG}_bench_other.json 2>> $OUT/${TAG}_bench.err; done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_light|k_shadow|k_ao_rays|k_taa|k_gbuffer|k_compose|refit|collapse" -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > $OUT/${TAG}_launches_bench.log 2>&1
i=0
for spec in "k_shadow_hints:k_shadow_hints:4" "k_shadow_rays_temporal:k_shadow_rays_temporal:4" "k_ao_rays_compact:k_ao_rays_compact:4" "k_light_shade:k_light_shade:4" "k_taa:k_taa:4"; do
  IFS=: read k key skip <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o $OUT/${TAG}_$key \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > $OUT/${TAG}_ncu_$key.log 2>&1
done
for tool in memcheck racecheck; do
  timeout 800 compute-sanitizer --tool $tool --log-file $OUT/${TAG}_sanitizer_$tool.log python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
      -k "default_scene_file_settings or synthetic_multi_light or taa_parity or tlas_refit or empty_and_degenerate or gbuffer_pass_culls or shadow_hints or temporal" > $OUT/${TAG}_sanitizer_${tool}_pytest.log 2>&1
  tail -1 $OUT/${TAG}_sanitizer_${tool}_pytest.log; tail -1 $OUT/${TAG}_sanitizer_$tool.log
done
tail -3 $OUT/${TAG}_pytest.log; tail -2 $OUT/${TAG}_pytest_exact_raygen.log; tail -1 $OUT/${TAG}_smoke.log; cut -c1-300 $OUT/${TAG}_bench.json
