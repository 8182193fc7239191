timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
AB_ENVS="LUZRT_ONE_VISIT=0 LUZRT_ONE_VISIT=1 LUZRT_RAY_KERNEL=plain" bash profiles/ab.sh ab10 build/variants/libluzrt_spec.so
AB_EXTRA="--shadow-type 2" AB_CONFIGS="c3" AB_ENVS="LUZRT_ONE_VISIT=0 LUZRT_ONE_VISIT=1 LUZRT_RAY_KERNEL=plain" bash profiles/ab.sh ab10b build/variants/libluzrt_spec.so
