#!/bin/bash
# Informational throughput lines for the five BASELINE.json configs on ONE GPU (the driver's metric is
# `python bench.py`).  Run on a GPU box:  gpurun -- tools/bench_configs.sh   -> gpurun_out/configs_*.json
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --steps 10 --warmup 3"
$B --T 10000                         2>&1 | tail -1 > gpurun_out/configs_c1_lgssm_n4_T1e4.json
$B                                   2>&1 | tail -1 > gpurun_out/configs_c1_lgssm_n4_T1e6.json
$B --nx 5 --ny 2                     2>&1 | tail -1 > gpurun_out/configs_lgssm_n5_T1e6.json
$B --nx 8 --ny 4                     2>&1 | tail -1 > gpurun_out/configs_c4_lgssm_n8_T1e6.json
$B --nx 8 --ny 4 --T 10000000 --steps 5 2>&1 | tail -1 > gpurun_out/configs_c4_lgssm_n8_T1e7.json
$B --no-host-model                   2>&1 | tail -1 > gpurun_out/configs_lgssm_n4_T1e6_timevarying_path.json
for lin in extended cubature gauss_hermite; do
  python bench.py --workload bearings --lin $lin --steps 9 --warmup 2 2>&1 | tail -1 > gpurun_out/configs_c2_bearings_$lin.json
done
python bench.py --workload bearings --lin extended --T 10000 --runs 100 --iters 20 --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/configs_c5_bearings_100runs_T1e4.json
python bench.py --workload bearings --lin extended --T 10000 --runs 100 --iters 20 --batched --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/configs_c5_bearings_100runs_T1e4_batched.json
python bench.py --workload bearings --lin unscented --steps 9 --warmup 2 2>&1 | tail -1 > gpurun_out/configs_c2_bearings_unscented.json
for lin in extended cubature gauss_hermite; do
  python bench.py --workload bearings --grad --lin $lin --steps 5 --warmup 2 2>&1 | tail -1 > gpurun_out/configs_c3_grad_$lin.json
done
python - <<'PY'
import glob, json
for p in sorted(glob.glob("gpurun_out/configs_*.json")):
    try:
        d = json.loads(open(p).read())
    except Exception as e:
        print(p, "ERR", open(p).read()[-300:]); continue
    if "roofline" in d:
        ns = d["roofline"]["north_star"]
        print(f'{p}: {d["ms_per_step"]:.3f} ms/pass  {d["value"]:.3e} steps/s  north-star frac {ns["frac_of_slower_bound"]:.3f}  stages {d["roofline"]["stage_ms"]}')
    else:
        print(f'{p}: {d}')
PY
