"""Selected raw metrics + a per-phase instruction breakdown of every launch in a .ncu-rep -> text.
usage: python tools/ncu_summary.py gpurun_out/<file>.ncu-rep "<title>" > profiles/<name>_ncu_summary.txt"""
import collections, csv, io, subprocess, sys
rep, title = sys.argv[1], sys.argv[2]
PATS = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct",
        "launch__registers_per_thread ", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max ", "smsp__inst_executed.sum ",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__shared_mem_per_block_dynamic", "smsp__pcsamp_warps_issue_stalled"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
print(f"# {title}; ncu --set full --clock-control none --import-source on; 1 x B200")
for li, r in enumerate(data):
    print("== " + r[4][:110] + " grid " + r[8] + " block " + r[7])
    for h, u, v in zip(hdr, units, r):
        if any((h + " ").startswith(p) or p.strip() == h or (p.endswith("stalled") and p in h and "not_issued" not in h) for p in PATS) and v not in ("", "0"):
            print(f"  {h} [{u}] = {v}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(li), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) < 3:
        continue
    sh = srows[1]
    ia, ie = sh.index("Source"), sh.index("Instructions Executed")
    ops = collections.Counter()
    for s in srows[2:]:
        if len(s) <= ie:
            continue
        t = s[ia].split()
        if not t:
            continue
        op = (t[1] if s[ia].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
        ops[op] += int(s[ie] or 0)
    tot = sum(ops.values())
    print("  executed warp instructions by opcode: " + ", ".join(f"{o} {100 * n / tot:.1f}%" for o, n in ops.most_common(14)) + f"  (total {tot})")
