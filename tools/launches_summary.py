"""Kernel launch list of one bench run under `ncu --metrics gpu__time_duration.sum` (tools/ncu_r02.sh launches) ->
profiles/r02_bench_launches_summary.txt: per kernel name: launches, total time, share.  Times under ncu are cold-cache and
serialised: only the SHARES are meaningful."""
import collections, csv, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r02_bench_launches.csv")
rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) >= 15 and r[0] != "ID"]
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    if r[12] != "gpu__time_duration.sum":
        continue
    v = float(r[14].replace(",", ""))
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[r[13]]
    tot[r[4]] += ms
    cnt[r[4]] += 1
allms = sum(tot.values())
lines = ["# r02 (final kernels): kernel launches of `python bench.py --steps 2 --warmup 1 --no-check --no-configs --sgld-steps 100` under "
         "ncu --metrics gpu__time_duration.sum --clock-control none (cold, serialised; shares only; first 400 launches)",
         "kernel | launches | total ms | share"]
for k, v in tot.most_common():
    lines.append(f"{k[:90]} | {cnt[k]} | {v:.3f} | {100 * v / allms:.2f}%")
out = os.path.join(ROOT, "profiles", "r02_bench_launches_summary.txt")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:14]))
