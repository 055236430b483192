#!/bin/bash
# launch list + full capture of the nx = 8 sub-warp pass; only the text summaries travel back (the report is > 64 MiB)
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-secondary --steps 1 --warmup 1 --nx 8 --ny 4"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/coop_launches_n8.csv $B > gpurun_out/coop_ncu1.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'k_coop_|k_chunk_|k_unit_' -c 7 -o /tmp/coop_n8_full -f $B > gpurun_out/coop_ncu2.log 2>&1
ncu -i /tmp/coop_n8_full.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/coop_n8_full_summary.txt
# hottest SASS lines (stall samples) of the three step-loop kernels
for k in k_coop_filter_reduce k_coop_filter_apply k_coop_smooth_apply; do
  ncu -i /tmp/coop_n8_full.ncu-rep --page source --csv -k regex:$k 2>/dev/null > /tmp/src_$k.csv
  python - $k <<'PY'
import csv, sys
k = sys.argv[1]
rows = list(csv.reader(open(f"/tmp/src_{k}.csv")))
hdr = None
for i, r in enumerate(rows):
    if "Source" in r and any("Sampl" in c for c in r):
        hdr = i; break
if hdr is None:
    print(k, "no source page", len(rows)); sys.exit(0)
h = rows[hdr]
si = h.index("Source")
samp = [i for i, c in enumerate(h) if c.startswith("# Samples") or c == "Samples" or "Sampling Data (All)" in c]
ci = samp[0] if samp else None
out = []
tot = 0
for r in rows[hdr + 1:]:
    try:
        v = float(r[ci].replace(",", "")) if ci is not None and r[ci] else 0.0
    except Exception:
        v = 0.0
    tot += v
    out.append((v, r[si][:110]))
# opcode histogram weighted by samples
import collections
hist = collections.Counter()
cnt = collections.Counter()
for v, s in out:
    op = s.split()[0] if s.split() else "?"
    if op.startswith("@"):
        op = s.split()[1] if len(s.split()) > 1 else op
    op = op.split(".")[0]
    hist[op] += v
    cnt[op] += 1
with open(f"gpurun_out/coop_src_{k}.txt", "w") as f:
    f.write(f"{k}: {len(out)} SASS lines, {tot:.0f} samples; header cols: {h}\n")
    f.write("opcode: samples share, static count\n")
    for op, v in hist.most_common(25):
        f.write(f"  {op:12s} {100 * v / max(tot, 1):5.1f} %  {cnt[op]}\n")
    f.write("hottest lines\n")
    for v, s in sorted(out, reverse=True)[:40]:
        f.write(f"  {100 * v / max(tot, 1):5.2f} %  {s}\n")
PY
done
ls gpurun_out | tail -5
