# 8-GPU runs of the bench (strong scaling of one frame): usage  gpurun --gpus 8 -- 'bash profiles/tools/scale8.sh <tag>'
TAG=${1:-r1}; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/${TAG}_smi8.txt 2>&1
run() { n=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n "$@"; }
run 8 --steps 50 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_c3_n8.json 2> $OUT/${TAG}_c3_n8.err
run 8 --config c5 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_c5_n8.json 2> $OUT/${TAG}_c5_n8.err
run 2 --steps 30 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_c3_n2.json 2> $OUT/${TAG}_c3_n2.err
run 4 --steps 30 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_c3_n4.json 2> $OUT/${TAG}_c3_n4.err
python - <<PY
import json
for f in ("c3_n2","c3_n4","c3_n8","c5_n8"):
    try:
        d=json.loads(open("$OUT/${TAG}_%s.json"%f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f value %.0f e2e %s kernels %s" % (d["ms_per_step"], d["value"], d.get("e2e",{}).get("ms_per_step"), {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["kernels_ms"].items()}))
    except Exception as e: print(f, "ERR", e)
PY
