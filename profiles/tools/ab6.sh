timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
AB_ENVS="LUZRT_RAY_PARTS=1 LUZRT_RAY_PARTS=2" bash profiles/ab.sh ab6 build/variants/libluzrt_parts.so
