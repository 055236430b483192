"""Resource usage and SASS opcode census of the hot kernels of libpsqrt.so -> profiles/ (run in the build container:
cuobjdump works without a GPU).  Shows what the sm_100a code is made of: DFMA / MUFU.RSQ64H for the algebra, SHFL for
the sub-warp reflectors, LDGSTS (cp.async) for the prefetch rings, UCGABAR_* / cluster barriers and DSMEM accesses for
the cluster mid scan."""
import collections
import re
import subprocess
import sys

SO = sys.argv[1] if len(sys.argv) > 1 else "sqrt-parallel-smoothers_b200/psqrt/libpsqrt.so"
WANT = [r"k_filter_reduceILi4ELi2ENS_6SrcValILi4ELi2E", r"k_filter_applyILi4ELi2ELb1ELb0ENS_6SrcValILi4ELi2EEENS_7WarpOutILi4ELi2E",
        r"k_smooth_applyILi4ENS_7SrcValTILi4EEENS_7WarpOutILi4ELi2E", r"k_mid_scan3INS_6CoopF2ILi4E", r"k_mid_scan3INS_6CoopF2ILi8E",
        r"k_coopr_filter_reduceILi8ELi4ELi4E", r"k_coopr_filter_applyILi8ELi4ELi4ELb0E", r"k_coopr_smooth_applyILi8ELi4E",
        r"k_unit_scanINS_6CoopF2ILi8E", r"k_unit_scanINS_6CoopS2ILi8E", r"k_chunk_startILi8ELb1E", r"k_chunk_endILi8E",
        r"k_filter_reduceILi5ELi2ENS_11SrcFusedCTB", r"k_carry_scanINS_6CoopF2ILi4E", r"k_adj_gradILi5ELi2E", r"k_applyILi5ELb1ELb1E"]
res = subprocess.run(["cuobjdump", "--dump-resource-usage", SO], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1)
    elif cur and "REG:" in line:
        usage[cur] = line.strip()
        cur = None
sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", sass)
census = {}
for b in blocks[1:]:
    name = b.split("\n", 1)[0].strip()
    ops = collections.Counter()
    for ln in b.splitlines():
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            ops[m.group(1).split(".")[0] if not m.group(1).startswith(("MUFU", "UCGABAR", "LDGSTS", "SHFL", "BAR", "MEMBAR", "ATOM", "RED", "LDS", "STS", "LD", "ST", "SYNCS", "CCTL", "ACQBULK", "UBLKCP")) else m.group(1)] += 1
    census[name] = ops
for pat in WANT:
    names = [n for n in usage if re.search(pat, n)]
    for n in names[:1]:
        dem = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
        print(dem[:150])
        print("   ", usage[n])
        ops = census.get(n, {})
        tot = sum(ops.values())
        keys = ["DFMA", "DMUL", "DADD", "MUFU.RSQ64H", "MUFU.RCP64H", "SHFL.IDX", "SHFL.UP", "SHFL.DOWN", "LDGSTS.E.64", "LDGSTS.E.BYPASS.128",
                "LDS.128", "LDS.64", "STS.128", "STS.64", "LDG.E.64", "LDG.E.128", "STG.E.64", "STG.E.128", "LDL.64", "STL.64", "LDL.128", "STL.128",
                "UCGABAR_ARV", "UCGABAR_WAIT", "BAR.SYNC", "MEMBAR.ALL.GPU", "MEMBAR.SC.SYS", "ATOMG.E.ADD.STRONG.GPU"]
        shown = {k: v for k, v in ops.items() if any(k.startswith(p) for p in keys) and v}
        top = ", ".join(f"{k} {v}" for k, v in sorted(shown.items(), key=lambda kv: -kv[1]))
        print(f"    SASS instructions {tot}: {top}")
        other = [k for k in ops if k.startswith(("UCGABAR", "LDGSTS", "SYNCS", "UBLKCP", "CCTL"))]
        if other:
            print("    " + ", ".join(f"{k} {ops[k]}" for k in other))
        print()
