mkdir -p gpurun_out
: > gpurun_out/r1f_f4.jsonl
for a in "--config c3 --volumetric 1" "--config c3 --shadow-type 2" "--config c3 --volumetric 2" "--config c2 --shadow-type 2 --volumetric 1" "--config c4 --shadow-type 2 --steps 5"; do
  echo "## $a" >> gpurun_out/r1f_f4.jsonl
  timeout 400 python bench.py $a --no-cpu-baseline --no-e2e --steps 10 --warmup 3 >> gpurun_out/r1f_f4.jsonl 2>> gpurun_out/r1f_f4.err
done
python - <<'PY'
import json
for l in open("gpurun_out/r1f_f4.jsonl"):
    if l.startswith("##"): print(l.strip(), end="  ")
    elif l.startswith("{"):
        d = json.loads(l); print("ms/step %.3f" % d["ms_per_step"], {k: (round(v,3) if isinstance(v,float) else v) for k,v in d["kernels_ms"].items() if k != "light_per_rank"})
PY
tail -5 gpurun_out/r1f_f4.err
