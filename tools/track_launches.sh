# per-launch times of the tracker's frame step (ncu, cold caches and serialised: shares, not bench values)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"preprocess|clip|scan_|icp_|beam" -c 12 --csv \
    --log-file gpurun_out/track_launches.csv python bench.py --workload track --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/track_launches.csv")) if len(r) > 5]
h = rows[0]
for r in rows[-6:]:
    print(r[h.index("Kernel Name")][:60], r[h.index("Metric Value")], r[h.index("Metric Unit")])
PY
