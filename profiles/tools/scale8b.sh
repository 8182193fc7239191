TAG=${1:-r1}; OUT=gpurun_out; mkdir -p $OUT
run() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n "$@"; }
: > $OUT/${TAG}_n8_ab.jsonl
for env in "LUZRT_RAY_PARTS=1 LUZRT_BAND_ROWS_MIN=48" "LUZRT_RAY_PARTS=2 LUZRT_BAND_ROWS_MIN=48" "LUZRT_RAY_PARTS=2 LUZRT_BAND_ROWS_MIN=90" "LUZRT_RAY_PARTS=2 LUZRT_BAND_ROWS_MIN=24"; do
  echo "## $env" >> $OUT/${TAG}_n8_ab.jsonl
  env $env bash -c "$(declare -f run); run 8 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e" >> $OUT/${TAG}_n8_ab.jsonl 2>> $OUT/${TAG}_n8_ab.err
done
python - <<PY
import json
for l in open("$OUT/${TAG}_n8_ab.jsonl"):
    if l.startswith("##"): print(l.strip(), end="  ")
    elif l.startswith("{"):
        d=json.loads(l); k=d["kernels_ms"]; print("ms/step %.3f light %.3f rays %.3f shade %.3f taa %.3f gather %.3f per-rank %s" % (d["ms_per_step"], k["light"], k["light_rays"], k["light_shade"], k["taa"], k["gather"], k["light_per_rank"]))
PY
