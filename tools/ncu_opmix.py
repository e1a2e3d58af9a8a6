"""Opcode mix and stall samples of one kernel from an ncu report's source page:
    ncu -i rep --page source --csv --kernel-name regex:NAME --launch-count 1 > src.csv ; python tools/ncu_opmix.py src.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
ix = {k: i for i, k in enumerate(rows[h])}
ops, samp = collections.Counter(), collections.Counter()
tot = tots = 0.0
first = None
for r in rows[h + 1:]:
    if len(r) < len(ix) or r[0] == "Address":
        continue
    src = r[ix["Source"]].split()
    op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
    n, s = float(r[ix["Instructions Executed"]]), float(r[ix["# Samples"]])
    first = first or n
    ops[op] += n; samp[op] += s; tot += n; tots += s
print(f"warp instructions per warp: {tot / first:.0f}   (warps {first:.0f}, samples {tots:.0f})")
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
    print(f"{op:10s} {n / first:8.1f} per warp   {100 * n / tot:5.1f} % of instructions   {100 * samp[op] / max(tots, 1):5.1f} % of stall samples")
