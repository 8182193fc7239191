for w in 8; do
python profiles/tools/rank_timeline.py $w 12 | tail -1
LUZRT_RAY_OVERLAP=0 python profiles/tools/rank_timeline.py $w 12 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_' --csv --log-file gpurun_out/rank8_launches.csv python profiles/tools/rank_timeline.py $w 8 > /dev/null 2>&1
done
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/rank8_launches.csv')) if len(r)>5 and r[0].isdigit()]
# last frame kernels
names=[(r[4].split('(')[0][:60], float(r[-1])) for r in rows]
for n,v in names[-14:]: print(n, v)
PY
