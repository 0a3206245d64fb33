# per-kernel durations of the window stage (ncu launch list, 1,000 cases)
cd /root/repo
mkdir -p gpurun_out/r2b
NC=1000 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k3_" --csv --log-file gpurun_out/r2b/k3_launches.csv python profiles/k3_variants.py > gpurun_out/r2b/k3_launch_run.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2b/k3_launches.csv')) if len(r)>5]
h=rows[0]; k=h.index('Kernel Name'); v=h.index('Metric Value')
for r in rows[1:16]: print(r[k][:40], r[v])
PY
