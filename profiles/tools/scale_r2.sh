# usage: bash profiles/tools/scale_r2.sh <tag> <n> [extra bench args]   (under gpurun --gpus n)
TAG=${1:-r2}; N=${2:-8}; shift; shift; OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline "$@" > $OUT/${TAG}_c3_n$N.json 2> $OUT/${TAG}_c3_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_c3_n$N.json").read().strip().splitlines()[-1])
    print("c3 n=$N ms/step %.3f value %.0f e2e %s kernels %s" % (d["ms_per_step"], d["value"], d.get("e2e",{}).get("ms_per_step"), {k:(round(v,3) if isinstance(v,float) else [round(x,2) for x in v]) for k,v in d["kernels_ms"].items()}))
    print("parity", {k: d.get("parity",{}).get(k) for k in ("agree","max_abs","psnr_db","pass","rows")}, "numa", d.get("e2e",{}).get("host_numa_binding"))
except Exception as e: print("ERR", e); print(open("$OUT/${TAG}_c3_n$N.err").read()[-1500:])
PY
