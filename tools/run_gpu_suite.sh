#!/bin/bash
# full GPU test suite + smoke + the default bench line (what the driver runs at round end)
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/gpu_suite.txt
tail -5 gpurun_out/gpu_suite.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['north_star']['frac_of_slower_bound'], d['secondary']['c4_strong']['ms_per_step'], d['cpu_baseline']['value'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bearings_ieks.csv python bench.py --workload bearings --lin extended --steps 1 --warmup 1 > gpurun_out/launches_bearings.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_bearings_ieks.csv')))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
seen = [(r[ki][:60], float(r[vi].replace(',', '')) / 1000) for r in rows[hdr + 1:] if len(r) > vi]
tail = seen[-40:]
for k, v in tail: print(f"{v:8.1f} us  {k}")
PY
