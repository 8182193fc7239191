timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash profiles/ab.sh ab7 build/variants/libluzrt_base.so build/variants/libluzrt_pluecker.so
