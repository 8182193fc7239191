timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
AB_ENVS="LUZRT_SHADOW_HINTS=0 LUZRT_SHADOW_HINTS=1" bash profiles/ab.sh ab11 build/variants/libluzrt_hints.so
