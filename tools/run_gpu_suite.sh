#!/bin/bash
# full GPU test suite + smoke + the default bench line (what the driver runs at round end)
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/gpu_suite.txt
tail -5 gpurun_out/gpu_suite.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['north_star']['frac_of_slower_bound'], d['secondary']['c4_strong']['ms_per_step'], d['cpu_baseline']['value'])"
