#!/bin/bash
mkdir -p gpurun_out
for t in 37888 18944; do
  export PSQRT_TARGET_CHUNKS=$t
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bearings_t$t.csv python bench.py --workload bearings --lin extended --steps 1 --warmup 1 > /dev/null 2>&1
  python - $t <<'PY'
import csv, sys, collections
t = sys.argv[1]
rows = list(csv.reader(open(f'gpurun_out/launches_bearings_t{t}.csv')))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
seen = [(r[ki], float(r[vi].replace(',', '')) / 1000) for r in rows[hdr + 1:] if len(r) > vi]
acc = collections.defaultdict(list)
for k, v in seen:
    for name in ('k_filter_reduce', 'k_mid_scan3', 'k_mid_scan2', 'k_filter_apply', 'k_smooth_apply', 'k_fused_trig'):
        if name in k:
            acc[name].append(v)
print('target', t, {k: (round(sum(v[-8:]) / len(v[-8:]), 1), len(v)) for k, v in acc.items()}, 'sum per iteration', round(sum(sum(v[-8:]) / len(v[-8:]) for v in acc.values()), 1))
PY
done
