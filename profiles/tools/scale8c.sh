TAG=${1:-r1}; OUT=gpurun_out; mkdir -p $OUT
run() { n=$1; shift; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n "$@"; }
run 8 --steps 50 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_c3_n8.json 2> $OUT/${TAG}_c3_n8.err
run 8 --config c5 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_c5_n8.json 2> $OUT/${TAG}_c5_n8.err
python - <<PY
import json
for f in ("c3_n8","c5_n8"):
    try:
        d=json.loads(open("$OUT/${TAG}_%s.json"%f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f value %.0f e2e %s kernels %s" % (d["ms_per_step"], d["value"], d.get("e2e",{}).get("ms_per_step"), {k:(round(v,3) if isinstance(v,float) else [round(x,2) for x in v]) for k,v in d["kernels_ms"].items()}))
    except Exception as e: print(f, "ERR", e)
PY
