#!/bin/bash
# nx=8 per-thread sweeps: chunk-length sweep (fewer resident threads -> smaller spill working set) + one ncu capture
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-secondary --steps 6 --warmup 3 --nx 8 --ny 4"
for c in 0 54 108 216; do
  $B --chunk $c 2>&1 | tail -1 > gpurun_out/exp_n8_chunk$c.json
done
$B --nx 5 --ny 2 --chunk 54 2>&1 | tail -1 > gpurun_out/exp_n5_chunk54.json
$B --nx 5 --ny 2 --chunk 108 2>&1 | tail -1 > gpurun_out/exp_n5_chunk108.json
python - <<'PY'
import glob, json
for p in sorted(glob.glob("gpurun_out/exp_n*_chunk*.json")):
    try:
        d = json.loads(open(p).read())
        print(p, f'{d["ms_per_step"]:.3f} ms', d["roofline"]["stage_ms"], d["config"]["chunk_len"])
    except Exception as e:
        print(p, "ERR", open(p).read()[-300:])
PY
timeout 600 ncu --set full --clock-control none -k regex:'k_filter_reduce|k_filter_apply|k_smooth_apply' -c 3 -o gpurun_out/exp_n8_full -f python bench.py --no-cpu-baseline --no-secondary --steps 1 --warmup 1 --nx 8 --ny 4 > gpurun_out/exp_n8_ncu.log 2>&1
tail -3 gpurun_out/exp_n8_ncu.log
