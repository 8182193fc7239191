timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
AB_ENVS="LUZRT_LIGHT_MINB=6 LUZRT_LIGHT_MINB=5" bash profiles/ab.sh ab5 build/variants/libluzrt_hemi.so build/variants/libluzrt_stream4.so build/variants/libluzrt_stream8.so
