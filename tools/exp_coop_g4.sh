#!/bin/bash
# two rows per lane (G = 4) against one row per lane (G = 8), same library (devlibs/libpsqrt_v5.so)
mkdir -p gpurun_out
export PSQRT_LIB=$PWD/devlibs/libpsqrt_v5.so
for m in 1 2 4 7; do PSQRT_COOP=$m PSQRT_COOP_G=4 timeout 300 python tools/check_coop.py 2>&1 | grep -E "WORST|rror|n=8 ny=4 T=1100" | tail -3; done
B="python bench.py --no-cpu-baseline --no-secondary --steps 6 --warmup 3 --nx 8 --ny 4"
for g in 4 8; do
  PSQRT_COOP_G=$g $B 2>&1 | tail -1 > gpurun_out/coopg${g}_n8.json
  PSQRT_COOP_G=$g $B --T 10000000 --steps 3 2>&1 | tail -1 > gpurun_out/coopg${g}_n8_T1e7.json
done
python - <<'PY'
import glob, json
for p in sorted(glob.glob("gpurun_out/coopg*.json")):
    try:
        d = json.loads(open(p).read())
        print(p, f'{d["ms_per_step"]:.3f} ms', d["roofline"]["stage_ms"], d["config"]["chunk_len"], round(d["roofline"]["north_star"]["frac_of_slower_bound"], 4))
    except Exception as e:
        print(p, "ERR", open(p).read()[-400:])
PY
