"""From the .ncu-rep files of a round-2 capture (tools/ncu_r02.sh) write
  * profiles/r02_<name>_ncu_summary.txt   -- selected raw metrics of the captured launch
  * profiles/r02_ncu_records.json         -- {key: {dram_bytes, tensor_pipe_active_pct, duration_ms, source}} read by bench.py
usage: python tools/ncu_records.py            (reads gpurun_out/r02_fwd_tc.ncu-rep, r02_fwd_tcl.ncu-rep, r02_sgld_step.ncu-rep)"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__pipe_tensor",
        "sm__throughput.avg.pct", "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct", "launch__registers_per_thread",
        "dram__throughput.avg.pct_of_peak", "sm__inst_executed_pipe_fma", "sm__pipe_fma_cycles_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "smsp__warp_issue_stalled", "lts__t_sector_hit_rate"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


def to_bytes(v, unit):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit)
    return v * f if (v is not None and f) else None


records = {}
for name, key, title in (("fwd_tc", "fwd_tc_kernel|S=10000|N=1000000|tf32_bf16", "fwd_tc_kernel at the bench configuration (python bench.py --steps 1 --warmup 1 --no-sgld --no-check --no-configs)"),
                         ("fwd_tcl", "fwd_tcl_kernel|S=50|N=100000|cfg5", "fwd_tcl_kernel, cfg 5 network, S=50 x N=100k (python tools/tcl_bench.py 50 100000)"),
                         ("sgld_step", "sgld_step_kernel|cfg5|B=128|50 steps", "sgld_step_kernel, cfg 5, batch 128, one launch of 50 steps (python tools/sgld_trace.py 128)")):
    rep = os.path.join(ROOT, "gpurun_out", f"r02_{name}.ncu-rep")
    if not os.path.exists(rep):
        print("missing", rep)
        continue
    hdr, units, rows = raw(rep)
    lines = [f"# r02: ncu --set full --clock-control none, {title}; 1 x B200"]
    for r in rows:
        lines.append("== " + r[4][:100] + " grid " + r[8] + " block " + r[7])
        d = {}
        for h, u, v in zip(hdr, units, r):
            if any(p in h for p in PATS):
                lines.append(f"  {h} [{u}] = {v}")
            d[h] = (num(v), u)
        rd = to_bytes(*d.get("dram__bytes_read.sum", (None, None)))
        wr = to_bytes(*d.get("dram__bytes_write.sum", (None, None)))
        dur, du = d.get("gpu__time_duration.sum", (None, None))
        dur_ms = dur * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(du, float("nan")) if dur is not None else None
        tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", (None, None))[0]
        if tp is None:
            for h in d:
                if h.startswith("sm__pipe_tensor") and "pct_of_peak_sustained_active" in h and d[h][0] is not None:
                    tp = d[h][0]
                    break
        records[key] = {"dram_bytes": (rd + wr) if (rd is not None and wr is not None) else None, "dram_bytes_read": rd,
                        "dram_bytes_write": wr, "tensor_pipe_active_pct": tp, "duration_ms_under_ncu": dur_ms,
                        "source": f"profiles/r02_{name}_ncu_summary.txt"}
    open(os.path.join(ROOT, "profiles", f"r02_{name}_ncu_summary.txt"), "w").write("\n".join(lines) + "\n")
# A later, cheap metrics-only pass over the same bench command (tools/ncu_r02.sh traffic) supersedes the traffic figures of the
# full-set capture when the kernel changed after it (r02: pacing of the CTA pairs went in after the full capture).
tcsv = os.path.join(ROOT, "gpurun_out", "r02_fwd_tc_traffic.csv")
key = "fwd_tc_kernel|S=10000|N=1000000|tf32_bf16"
if os.path.exists(tcsv) and key in records:
    rows = [r for r in csv.reader(open(tcsv)) if len(r) > 6]
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    vals = {}
    for r in rows[1:]:
        if "fwd_tc_kernel" in r[col["Kernel Name"]]:
            vals[r[col["Metric Name"]]] = (num(r[col["Metric Value"]]), r[col["Metric Unit"]])
    rd, wr = to_bytes(*vals.get("dram__bytes_read.sum", (None, None))), to_bytes(*vals.get("dram__bytes_write.sum", (None, None)))
    if rd is not None and wr is not None:
        rec = records[key]
        rec["full_set_capture_before_pacing"] = {"dram_bytes": rec["dram_bytes"], "dram_bytes_read": rec["dram_bytes_read"],
                                                 "dram_bytes_write": rec["dram_bytes_write"]}
        rec.update({"dram_bytes": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
                    "l2_hit_pct": vals.get("lts__t_sector_hit_rate.pct", (None,))[0],
                    "source": "profiles/r02_fwd_tc_traffic.csv (metrics pass of the shipped kernel); tensor pipe: profiles/r02_fwd_tc_ncu_summary.txt"})
        tp = vals.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", (None,))[0]
        if tp is not None:
            rec["tensor_pipe_active_pct"] = tp
        open(os.path.join(ROOT, "profiles", "r02_fwd_tc_traffic.csv"), "w").write(open(tcsv).read())
json.dump(records, open(os.path.join(ROOT, "profiles", "r02_ncu_records.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(records, indent=1))
