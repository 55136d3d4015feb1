"""Aggregate an `ncu --page source --csv` dump (SASS view): executed instructions / stall samples per opcode,
and the hottest instructions.  python tools/ncu_source.py src.csv [topN]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
by_op = collections.Counter(); samp_op = collections.Counter(); wf = collections.Counter()
items = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    n = int(float(r[ix["Instructions Executed"]] or 0)); s = int(float(r[ix["# Samples"]] or 0))
    w = int(float(r[ix["L1 Wavefronts Shared"]] or 0))
    by_op[op] += n; samp_op[op] += s; wf[op] += w
    items.append((s, n, w, r[ix["Address"]], src))
tot = sum(by_op.values()); tots = sum(samp_op.values())
print(f"total executed warp-instructions {tot:,}; samples {tots:,}")
for op, n in by_op.most_common(18):
    print(f"  {op:28s} {n:>13,} ({100*n/tot:5.1f}%)  samples {100*samp_op[op]/max(tots,1):5.1f}%  smem wavefronts {wf[op]:,}")
print("hottest by samples:")
for s, n, w, a, src in sorted(items, reverse=True)[:top]:
    print(f"  {s:7d} {n:>11,} wf={w:<10,} {a} {src[:90]}")
