AB_CONFIGS="c3 c2" AB_ENVS="LUZRT_RAY_PARTS=2 LUZRT_RAY_PARTS=1" bash profiles/ab.sh ab9 build/variants/libluzrt_pluecker.so
