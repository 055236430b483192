#!/bin/bash
# chunk-count target sweep (PSQRT_TARGET_CHUNKS): nx = 5 (bearings T = 1e5, LGSSM T = 1e6), nx = 4 T = 1e6
mkdir -p gpurun_out
for t in 37888 28416 18944 14208 9472; do
  export PSQRT_TARGET_CHUNKS=$t
  for rep in 1 2; do
    python bench.py --workload bearings --lin extended --steps 8 --warmup 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bearings T=1e5 target $t:', round(d['ms_per_call'],3))"
  done
  python bench.py --no-cpu-baseline --no-secondary --steps 10 --warmup 3 --nx 5 --ny 2 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lgssm nx5 T=1e6 target $t:', round(d['ms_per_step'],4), d['config']['chunk_len'])"
  python bench.py --no-cpu-baseline --no-secondary --steps 20 --warmup 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lgssm nx4 T=1e6 target $t:', round(d['ms_per_step'],4), d['config']['chunk_len'])"
done
