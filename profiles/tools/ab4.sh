timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash profiles/ab.sh ab4 build/variants/libluzrt_split.so build/variants/libluzrt_fastgen.so build/variants/libluzrt_hemi.so
timeout 300 python bench.py --config c5 --no-e2e --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/ab4_c5.json 2>gpurun_out/ab4_c5.err; python profiles/bench_summary.py gpurun_out/ab4_c5.json 2>/dev/null | tail -3
