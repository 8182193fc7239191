timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
AB_CONFIGS="c3 c2" bash profiles/ab.sh ab8a build/variants/libluzrt_pluecker.so build/variants/libluzrt_onepass_k3.so
AB_CONFIGS="c3 c2" AB_ENVS="LUZRT_LIGHT_MINB=5 LUZRT_LIGHT_MINB=7 LUZRT_RAY_PARTS=2" bash profiles/ab.sh ab8b build/variants/libluzrt_onepass.so
bash profiles/ab.sh ab8c build/variants/libluzrt_onepass.so
