python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_parity.py -m gpu -x -q 2>&1 | tail -3
AB_ENVS="X=1" AB_CONFIGS="c3 c4 c2 c5" bash profiles/tools/ab_env.sh tilehints2
python profiles/tools/rank_timeline.py 8 12 | tail -1
