mkdir -p gpurun_out
P=$PWD/sqrt-parallel-smoothers_b200/psqrt
show='import sys,json
d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["stage_ms"], d["e2e"]["ms_per_step"])'
export PSQRT_LIB=$P/libpsqrt_dev.so
timeout 900 python -m pytest tests -m gpu -q -k "lgssm and 4-2-1000 or time_varying and 4-2-300 or by_value and 4-2-1000 or fake_ranks and 4-2-3 or full_size or batched or sequential_oracle or non_triangular" 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$show"
ncu --set full --clock-control none --import-source on -k regex:"k_filter_reduce|k_filter_apply|k_smooth_apply" -s 6 -c 3 -o gpurun_out/prof_v7 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
