#!/bin/bash
# A/B of sub-warp sweep builds: devlibs/libpsqrt_<v>.so (nx = 8 only)
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-secondary --steps 6 --warmup 3 --nx 8 --ny 4"
for v in "$@"; do
  export PSQRT_LIB=$PWD/devlibs/libpsqrt_$v.so
  PSQRT_COOP=7 timeout 300 python tools/check_coop.py 2>&1 | grep -E "WORST|rror" | tail -3
  PSQRT_COOP=7 $B 2>&1 | tail -1 > gpurun_out/coopv_${v}_n8.json
  PSQRT_COOP=7 $B --T 10000000 --steps 3 2>&1 | tail -1 > gpurun_out/coopv_${v}_n8_T1e7.json
done
python - <<'PY'
import glob, json
for p in sorted(glob.glob("gpurun_out/coopv_*.json")):
    try:
        d = json.loads(open(p).read())
        print(p, f'{d["ms_per_step"]:.3f} ms', d["roofline"]["stage_ms"], d["config"]["chunk_len"], round(d["roofline"]["north_star"]["frac_of_slower_bound"], 4))
    except Exception as e:
        print(p, "ERR", open(p).read()[-400:])
PY
