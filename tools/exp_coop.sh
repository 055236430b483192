#!/bin/bash
# parity of every mix of per-thread / sub-warp sweeps, then timings at nx = 8 and 6
mkdir -p gpurun_out
for m in 0 2 4 6 1 3 7; do PSQRT_COOP=$m timeout 300 python tools/check_coop.py 2>&1 | tail -12; done > gpurun_out/coop_check.txt 2>&1
grep -E "WORST|Error|error" gpurun_out/coop_check.txt | head -20
B="python bench.py --no-cpu-baseline --no-secondary --steps 6 --warmup 3"
for m in 0 7; do
  PSQRT_COOP=$m $B --nx 8 --ny 4 2>&1 | tail -1 > gpurun_out/coop_n8_mask$m.json
  PSQRT_COOP=$m $B --nx 6 --ny 4 2>&1 | tail -1 > gpurun_out/coop_n6_mask$m.json
done
python - <<'PY'
import glob, json
for p in sorted(glob.glob("gpurun_out/coop_n*_mask*.json")):
    try:
        d = json.loads(open(p).read())
        print(p, f'{d["ms_per_step"]:.3f} ms', d["roofline"]["stage_ms"], d["config"]["chunk_len"], d["roofline"]["north_star"]["frac_of_slower_bound"])
    except Exception as e:
        print(p, "ERR", open(p).read()[-400:])
PY
