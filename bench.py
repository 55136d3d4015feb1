#!/usr/bin/env python
"""Headline benchmark: posterior-predictive sample.point MLP evals/s (BASELINE.json metric).

Workload (BASELINE.json configs[3], "synthetic predictive sweep"): N = 1M points x S = 10k SGLD samples, 2x200 ReLU MLP
(D = 16).  The S = 10k samples are sharded over the GPUs (STRONG scaling: 1 GPU evaluates all 10k, 8 GPUs 1250 each); one
"step" is one predict_mv over the whole job: operand-image prep + fused forward + Welford reduction on every rank, then --
for more than one GPU -- bnn_moments_allreduce (NCCL reduce-scatter of float64 raw moments, slice finalisation, all-gather).

    python bench.py [--gpus N] [--steps K] [--warmup W]           # our CUDA path
    python bench.py --impl reference ...                          # the reference's CPU torch loop

Prints ONE JSON line (field notes: DESIGN.md "Measurement").  Secondary blocks on the same line: `configs` (cfg 1, 2, 3 and
cfg 5's predictive pass through the drop-in classes, 1 GPU only), `sgld` (cfg 5: fused SGLD steps/s, one chain per GPU),
`spot_check` (the timed result against the CPU oracle on a sample of points, outside the timed region).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS = [16, 200, 200, 1]
ACT = "relu"
N_POINTS = 1_000_000
S_TOTAL = 10_000
METRIC = "posterior-predictive sample.point MLP evals/sec"
UNIT = "evals/s"
FLOP_PER_EVAL = 2 * sum(i * o for i, o in zip(DIMS[:-1], DIMS[1:]))   # 86,800
SGLD_DIMS, SGLD_BATCH, SGLD_ROWS, SGLD_KEEP = [90, 512, 512, 512, 1], 128, 10_000_000, 50     # BASELINE.json configs[4]
SGLD_FLOP_PER_STEP = 3 * SGLD_BATCH * 2 * sum(i * o for i, o in zip(SGLD_DIMS[:-1], SGLD_DIMS[1:]))   # ~438 MFLOP (fwd + bwd)


def workload_name(n_gpus):
    return (f"cfg4 synthetic predictive sweep: N={N_POINTS} points x S={S_TOTAL} SGLD samples, "
            f"{DIMS[0]}-{DIMS[1]}-{DIMS[2]}-{DIMS[3]} {ACT} MLP, samples sharded over {n_gpus} GPU(s)")


def shard_range(S, rank, world):
    base, extra = divmod(int(S), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


# --------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the reference's own per-sample torch loop (oracle port)
# --------------------------------------------------------------------------------------
def cpu_nets(S, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    base = []
    for i, o in zip(DIMS[:-1], DIMS[1:]):
        bound = (6.0 / (i + o)) ** 0.5                          # xavier-uniform, util.py:19
        base.append(((torch.rand(o, i, generator=g) * 2 - 1) * bound, (torch.rand(o, generator=g) * 2 - 1) / i ** 0.5))
    return [[(W + 0.05 * torch.randn(W.shape, generator=g), b + 0.05 * torch.randn(b.shape, generator=g)) for W, b in base]
            for _ in range(S)]


def time_cpu_sample(S, N, reps=1):
    """Time predict_mv of the reference's loop (sample_predict + mean/var) on S x N evals."""
    import torch
    from oracle import bnn_oracle as O           # bench's cpu_baseline leg may execute the oracle
    torch.set_num_threads(os.cpu_count() or 1)
    nets = cpu_nets(S, 1)
    X = torch.randn(N, DIMS[0], generator=torch.Generator().manual_seed(2))
    best = float("inf")
    with torch.no_grad():
        O.sample_predict(nets[:1], X[:1000], ACT, nout=1)      # warm-up
        for _ in range(reps):
            t0 = time.perf_counter()
            pred = O.sample_predict(nets, X, ACT, nout=1)
            O.predict_mv(pred, N)
            best = min(best, time.perf_counter() - t0)
    return S * N / best, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    S, N = 8, 250_000                                            # bounded sample of the workload per step
    for _ in range(args.warmup):
        time_cpu_sample(2, 50_000)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        time_cpu_sample(S, N)
    dt = time.perf_counter() - t0
    value = args.steps * S * N / dt
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus), "sample_per_step": f"S={S} x N={N} of the workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x (S={S} samples x N={N} points), torch CPU, {cores} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# --------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.path = None, f"/tmp/bnn_clocks_{os.getpid()}.csv"
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            # median over the upper half = samples taken under load
            hi = sorted(sm)[len(sm) // 2:]
            out.update(sm_mhz=statistics.median(hi), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def ncu_record(key):
    """dram bytes / tensor-pipe activity of the dominant kernel from the committed ncu capture of THIS configuration
    (profiles/r02_ncu_records.json, written by tools/ncu_records.py from the .ncu-rep of a bench run); None when the
    configuration was never captured."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_records.json"))).get(key)
    except Exception:
        return None


def make_bank(arch, rank_of_shard, lo, hi, dev):
    """Shard [lo, hi) of the synthetic chain: xavier base + 0.05 N(0,1), one generator per shard."""
    import torch
    base = arch.pack(cpu_nets(1, 0)[:1])[0].to(dev)
    gw = torch.Generator(device=dev).manual_seed(100 + rank_of_shard)
    bank = (base[None, :] + 0.05 * torch.randn(hi - lo, arch.stride, generator=gw, device=dev)) * arch.valid_mask().to(dev)
    return bank.contiguous()


def run_ours(args):
    import torch
    import torch.distributed as dist
    from bnn_b200 import _lib, ops
    from bnn_b200.layout import Arch
    from bnn_b200.distributed import merge_shards

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    # ---- CPU legs: 1-GPU runs only, and before anything else is resident (at N > 1 the other ranks would otherwise spin
    #      inside NCCL while rank 0 times the host loop)
    cpu_base, sgld_cpu = None, None
    cores = os.cpu_count() or 1
    if world == 1:
        cpu_val, cpu_t = time_cpu_sample(16, 250_000)
        cpu_base = {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"S=16 samples x N=250000 points of the same workload, torch CPU loop, {cores} threads, {cpu_t:.1f} s"}
        if not args.no_sgld:
            try:
                sgld_cpu = sgld_cpu_baseline()
            except Exception as e:
                sgld_cpu = {"error": repr(e)[:200]}

    nccl_merge = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's banner must not share stdout with the JSON line
        dist.init_process_group("nccl", device_id=dev)
        from bnn_b200.distributed import NcclMoments
        nccl_merge = NcclMoments(dev)                             # the collective of the C ABI (bnn_moments_allreduce)
    n_gpus = world

    arch = Arch(DIMS, ACT)
    lo, hi = shard_range(S_TOTAL, rank, world)
    S = hi - lo
    g = torch.Generator(device=dev).manual_seed(1234 + 4)
    X = torch.randn(N_POINTS, DIMS[0], generator=g, device=dev)      # X ~ N(0,1), replicated
    bank = make_bank(arch, rank, lo, hi, dev)
    X_host = X.cpu().pin_memory()
    bank_host = bank.cpu().pin_memory()

    def step_device():
        out = ops.forward(arch, X, bank, want_pred=False, want_moments=(world == 1), want_partials=(world > 1),
                          precision=args.precision)
        if world > 1:
            return nccl_merge.merge(out["part_mean"], out["part_m2"], S)
        return out["mean"], out["var"]

    host_out = {}

    def step_e2e():
        # the C ABI's host-buffer entry point: H2D of X and of the bank (in chunks behind the evaluation), launches, D2H
        nonlocal host_out
        if world == 1:
            host_out = ops.forward_host(arch, X_host, bank_host, want_moments=True, precision=args.precision, out=host_out)
            return host_out["mean"], host_out["var"]
        host_out = ops.forward_host(arch, X_host, bank_host, want_moments=False, want_partials=True, precision=args.precision,
                                    out=host_out)
        mean, var = nccl_merge.merge(host_out["part_mean"].to(dev, non_blocking=True), host_out["part_m2"].to(dev, non_blocking=True), S)
        return mean.cpu(), var.cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            r = fn()
        e1.record()
        barrier()
        ms = max(e0.elapsed_time(e1), 0.0)
        wall_ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            t = torch.tensor([ms, wall_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall_ms = float(t[0]), float(t[1])
        return ms, wall_ms, r

    warm = max(args.warmup, 3)
    clocks = ClockSampler(local) if rank == 0 else None
    ms, _, (mean, var) = timed(step_device, args.steps, warm)
    clk = clocks.stop() if clocks else None
    e2e_steps = max(1, min(args.steps, 3))
    # the host path synchronises inside the library (its own stream): its time is wall time around the calls
    _, wall_e2e, (mean_h, var_h) = timed(step_e2e, e2e_steps, 1)
    e2e_match = float((mean_h.to(dev).reshape(-1) - mean.reshape(-1)).abs().max())

    # ---- spot check against the CPU oracle (outside the timed region): all S_TOTAL samples on a sample of points
    spot = None
    if rank == 0 and not args.no_check:
        try:
            spot = spot_check(arch, mean, var, X, dev, world)
        except Exception as e:
            spot = {"error": repr(e)[:200]}

    sgld = None
    if not args.no_sgld:
        try:
            sgld = sgld_bench(dev, rank, args.sgld_steps)
            if world > 1:
                t = torch.tensor([sgld["steps_per_s_per_chain"]], device=dev, dtype=torch.float64)
                dist.all_reduce(t)
                sgld["steps_per_s_all_chains"] = float(t)
            else:
                sgld["steps_per_s_all_chains"] = sgld["steps_per_s_per_chain"]
            sgld["chains"] = world
            try:
                sgld["chains_per_gpu"] = sgld_multi_chain_block(dev, rank)
            except Exception as e:
                sgld["chains_per_gpu"] = {"error": repr(e)[:300]}
            if sgld_cpu is not None:
                sgld["cpu"] = sgld_cpu
        except Exception as e:      # the secondary metric must never take the headline down
            sgld = {"error": repr(e)[:300]}

    configs = None
    precisions = None
    bbb_train = None
    dropout_train = None
    sgld_cfg1 = None
    if world == 1 and not args.no_configs:
        try:
            configs = config_blocks(dev)
        except Exception as e:
            configs = {"error": repr(e)[:300]}
        try:
            del mean_h, var_h
            torch.cuda.empty_cache()
            precisions = precision_blocks(dev)
        except Exception as e:
            precisions = {"error": repr(e)[:300]}
        try:
            bbb_train = bbb_train_block(dev)
        except Exception as e:
            bbb_train = {"error": repr(e)[:300]}
        try:
            sgld_cfg1 = sgld_cfg1_block(dev)
        except Exception as e:
            sgld_cfg1 = {"error": repr(e)[:300]}
        try:
            dropout_train = dropout_train_block(dev)
        except Exception as e:
            dropout_train = {"error": repr(e)[:300]}

    evals_per_step = float(S_TOTAL) * N_POINTS
    value = evals_per_step * args.steps / (ms * 1e-3)
    e2e_value = evals_per_step * e2e_steps / (wall_e2e * 1e-3)
    finite = bool(torch.isfinite(mean).all() and torch.isfinite(var).all())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bf16_peak, peak_src = (peaks.get("bf16_tflops_sustained"), "measured (MEASURED_PEAKS.json, sustained: the kernel runs for seconds)") \
            if peaks.get("bf16_tflops_sustained") else (1400.0, "fallback (B200_PROFILING.md sustained)")
        fp32_peak = float(_lib.lib().bnn_measure_fp32_tflops())
        per_gpu_tflops = FLOP_PER_EVAL * float(S) * N_POINTS * args.steps / (ms * 1e-3) / 1e12
        tc = args.precision in ("tf32_bf16", "tf32x3", "tf32", "bf16")
        units = {"tf32_bf16": 2 + 1 + 1, "tf32x3": 3 * 2, "tf32": 2, "bf16": 1}.get(args.precision, 0)
        scheme = {"tf32_bf16": "one TF32 pass + two BF16 correction passes", "tf32x3": "three TF32 passes",
                  "tf32": "one TF32 pass", "bf16": "one BF16 pass"}.get(args.precision, "")
        rec = ncu_record(f"fwd_tc_kernel|S={S}|N={N_POINTS}|{args.precision}") if tc else None
        alg_bytes = 4.0 * (S * arch.n_params + N_POINTS * DIMS[0] + 2 * N_POINTS * DIMS[-1])
        if tc:
            roof = {"bound": "tensor", "achieved": per_gpu_tflops, "peak": bf16_peak, "unit": "TFLOP/s",
                    "frac": per_gpu_tflops / bf16_peak,
                    "traffic": rec.get("dram_bytes") if rec else None, "traffic_unit": "bytes/launch",
                    "traffic_source": ("ncu capture of this configuration: " + rec.get("source", "")) if rec
                    else "no ncu capture of this configuration committed",
                    "algorithmic_bytes": alg_bytes,
                    "kernel": "fwd_tc_kernel (tcgen05 kind::tf32 + kind::f16, cta_group::2, A operand of layer 2 in TMEM)",
                    "flop_per_eval": FLOP_PER_EVAL, "peak_source": peak_src + "; dense bf16",
                    "note": f"achieved = ALGORITHMIC flop/s (2*sum in*out per eval).  {args.precision} issues {scheme} per algorithmic "
                            f"MAC = {units} bf16-equivalent units, so frac is capped at {1.0 / units:.3f}",
                    "issued_bf16_equiv_frac": units * per_gpu_tflops / bf16_peak,
                    "tensor_pipe_active_pct_ncu": rec.get("tensor_pipe_active_pct") if rec else None}
        else:
            roof = {"bound": "fp32", "achieved": per_gpu_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
                    "frac": per_gpu_tflops / fp32_peak if fp32_peak > 0 else None, "traffic": None,
                    "kernel": "fwd_simt_kernel<8>", "flop_per_eval": FLOP_PER_EVAL,
                    "peak_source": "FP32 FFMA microbenchmark measured live (bnn_measure_fp32_tflops); "
                                   "MEASURED_PEAKS.json has no FP32-SIMT figure",
                    "frac_of_bf16_tensor_peak": per_gpu_tflops / bf16_peak, "bf16_peak": bf16_peak}
        launches_per_step = (3 if tc else 2) + (3 if world > 1 else 0)   # + pack / finalise / unpack kernels around the two NCCL calls
        h2d = int((bank.numel() + X.numel()) * 4) + (int(2 * N_POINTS * 8) if world > 1 else 0)
        d2h = int(2 * N_POINTS * DIMS[-1] * 4) + (int(2 * N_POINTS * 8) if world > 1 else 0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None,
            "dtype": {"tf32_bf16": "tf32+bf16 (TF32 main term + two BF16 correction terms, FP32 accumulate; parity 1e-5)",
                      "tf32x3": "tf32x3 (3-pass error-compensated TF32, FP32 accumulate; parity 1e-5)",
                      "tf32": "tf32 (single pass, bound 5e-3)", "bf16": "bf16 (single pass, bound 2e-2)",
                      "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": {"workload": workload_name(n_gpus), "precision": args.precision, "samples_per_gpu": S,
                       "l2": f"inputs {(bank.numel() + X.numel()) * 4 / 1e6:.0f} MB per GPU > 126 MB L2 (no flush needed)",
                       "parallelism": f"samples sharded x{n_gpus}" + (", bnn_moments_allreduce: NCCL reduce-scatter of float64 raw moments + slice finalisation + all-gather" if n_gpus > 1 else "")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": wall_e2e / e2e_steps, "steps": e2e_steps,
                    "api": "bnn_forward_host (C ABI, host pointers: H2D of X and of the bank in chunks overlapped with the "
                           "evaluation, D2H of the result, per step)" + ("; shard moments re-uploaded for the NCCL merge" if world > 1 else ""),
                    "max_abs_diff_vs_device_resident_mean": e2e_match},
            "gpu_launches": args.steps * launches_per_step,
            "roofline": roof,
            "fp32_ffma_peak_tflops": fp32_peak,
            "cpu_baseline": cpu_base if cpu_base is not None else {"value": None, "unit": UNIT, "cores": cores, "kind": "port",
                                                                   "sample": "not timed at N > 1 (rank 0 at N = 1 only)"},
            "clocks": clk, "outputs_finite": finite, "spot_check": spot, "sgld": sgld, "configs": configs,
            "precision_modes": precisions, "bbb_train": bbb_train, "dropout_train": dropout_train, "sgld_cfg1": sgld_cfg1,
        }
        _emit(line)
    if world > 1:
        nccl_merge.close()
        dist.destroy_process_group()


def spot_check(arch, mean, var, X, dev, world, n_pts=256):
    """The timed result against the CPU oracle (the reference's loop) on n_pts points x all S_TOTAL samples."""
    import torch
    from oracle import bnn_oracle as O
    O.warmup()
    torch.set_num_threads(os.cpu_count() or 1)
    idx = torch.linspace(0, N_POINTS - 1, n_pts).long()
    Xs = X[idx.to(dev)].cpu()
    preds = []
    t0 = time.perf_counter()
    for r in range(world):
        lo, hi = shard_range(S_TOTAL, r, world)
        b = make_bank(arch, r, lo, hi, dev).cpu()
        nets = [arch.unpack(b[s]) for s in range(b.shape[0])]
        with torch.no_grad():
            preds.append(O.sample_predict(nets, Xs, ACT, nout=1))
    pred = torch.cat(preds)
    m_ref, v_ref = O.predict_mv(pred, n_pts)
    m, v = mean[idx.to(dev)].cpu().double(), var[idx.to(dev)].cpu().double()
    return {"points": n_pts, "samples": int(pred.shape[0]),
            "mean_max_norm_err": float((m - m_ref.double()).abs().max() / m_ref.abs().max()),
            "var_max_norm_err": float((v - v_ref.double()).abs().max() / v_ref.abs().max()),
            "tolerance": 1e-5, "oracle_s": time.perf_counter() - t0}


# --------------------------------------------------------------------------------------
def gpu_ms(fn, reps=20, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def precision_blocks(dev):
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    _precision_doc = """The other arithmetic modes of the resident tensor-core forward on a slice of cfg 4 (N = 1M points x 256 samples), each
    with the error it actually makes against the CPU oracle on 256 points -- the "stated looser bound" of the faster modes."""
    import torch
    from oracle import bnn_oracle as O
    O.warmup()
    arch = Arch(DIMS, ACT)
    Sb = 256
    g = torch.Generator(device=dev).manual_seed(99)
    Xb = torch.randn(N_POINTS, DIMS[0], generator=g, device=dev)
    bank = make_bank(arch, 0, 0, Sb, dev)
    idx = torch.linspace(0, N_POINTS - 1, 256).long()
    nets = [arch.unpack(b) for b in bank.cpu()]
    with torch.no_grad():
        ref = O.sample_predict(nets, Xb[idx.to(dev)].cpu(), ACT, nout=1)
    m_ref, v_ref = O.predict_mv(ref, 256)
    out = {}
    for prec, units in (("tf32_bf16", 4), ("tf32x3", 6), ("tf32", 2), ("fp32", 0)):
        box = {}

        def fn():
            box["r"] = ops.forward(arch, Xb, bank, precision=prec)
        ms = gpu_ms(fn, reps=5, warm=2)
        r = box["r"]
        m, v = r["mean"][idx.to(dev)].cpu().double().reshape(-1), r["var"][idx.to(dev)].cpu().double().reshape(-1)
        out[prec] = {"evals_per_s": Sb * N_POINTS / (ms * 1e-3), "ms": ms, "samples": Sb,
                     "tflops_algorithmic": FLOP_PER_EVAL * Sb * N_POINTS / (ms * 1e-3) / 1e12,
                     "bf16_equiv_units_per_mac": units or None,
                     "mean_max_norm_err": float((m - m_ref.double().reshape(-1)).abs().max() / m_ref.abs().max()),
                     "var_max_norm_err": float((v - v_ref.double().reshape(-1)).abs().max() / v_ref.abs().max()),
                     "kernel": "fwd_simt_kernel" if prec == "fp32" else "fwd_tc_kernel"}
    return out


def bbb_train_block(dev):
    """f2: the fused Bayes-by-Backprop training kernel at the cfg 2 shape (1-50-50-1 tanh, batch 32, 1000 rows): steps/s of one
    launch per epoch, next to the reference's recipe on the host cores (oracle port, torch CPU autograd + Adam)."""
    import torch
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    from oracle import bnn_oracle as O
    dims, B, N = [1, 50, 50, 1], 32, 1000
    arch = Arch(dims, "tanh")
    g = torch.Generator().manual_seed(0)
    X = torch.rand(N, 1, generator=g) * 6 - 3
    y = torch.sin(X)
    mus = [torch.empty(o, i + 1).uniform_(-0.3, 0.3, generator=g) for i, o in zip(dims[:-1], dims[1:])]
    rhos = [torch.full((o, i + 1), -5.0) for i, o in zip(dims[:-1], dims[1:])]
    tr = ops.BbbTrainer(arch, B, dev, seed=1)
    tr.set_state(arch.pack_bias_last(mus), arch.pack_bias_last(rhos))
    tr.bind(X.to(dev), y.to(dev))
    steps = (N + B - 1) // B
    perms = [torch.randperm(N, generator=g).to(dev) for _ in range(8)]
    for q in perms[:2]:
        tr.run(q, 0, steps)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 100
    e0.record()
    for r in range(reps):
        tr.run(perms[r % 8], 0, steps)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * steps)
    n_cpu = 100
    batches = [torch.randperm(N, generator=g)[:B] for _ in range(n_cpu)]
    eps = [[torch.randn(B, o) for o in dims[1:]] for _ in range(n_cpu)]
    cpu, cpu_threads = 0.0, 1
    for nt in (1, os.cpu_count() or 1):      # ops this small are often faster on one thread: report the better of the two
        torch.set_num_threads(nt)
        O.bbb_train_steps(mus, rhos, X, y, batches[:5], "tanh", eps[:5])
        t0 = time.perf_counter()
        O.bbb_train_steps(mus, rhos, X, y, batches, "tanh", eps)
        r = n_cpu / (time.perf_counter() - t0)
        if r > cpu:
            cpu, cpu_threads = r, nt
    return {"workload": "cfg2 shape: BNN_BBB 1-50-50-1 tanh, batch 32, 1000 rows, launches of one epoch (32 steps)",
            "kernel": "bbb_train_kernel (single-CTA persistent kernel: gather, local-reparameterisation forward, MSE, backward, "
                      "closed-form KL gradient, Adam; state in shared memory)",
            "steps_per_s": 1e6 / us, "us_per_step": us, "state_finite": bool(torch.isfinite(tr.mu).all()),
            "cpu": {"steps_per_s": cpu, "cores": cpu_threads, "kind": "port",
                    "sample": f"{n_cpu} steps, torch CPU autograd + Adam, best of 1 / all threads"}}


def sgld_multi_chain_block(dev, rank):
    """cfg 5 with SEVERAL independent chains on one GPU: each chain's launches on its own stream, CTAs per launch capped
    (bnn_sgld_set_grid_cap) so the chains' dependency stalls interleave.  Same step kernel, same launches of 50 steps; the data
    set is a 1M-row slice (the gather touches 6,400 random rows per launch either way)."""
    import torch
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    arch = Arch(SGLD_DIMS, "tanh")
    rows, n = 1_000_000, SGLD_KEEP
    g = torch.Generator(device=dev).manual_seed(17)
    X = torch.randn(rows, SGLD_DIMS[0], generator=g, device=dev)
    y = torch.randn(rows, 1, generator=g, device=dev)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    out = {}
    try:
        for chains in (2, 4):
            ops.sgld_set_grid_cap(sms // chains)
            streams = [torch.cuda.Stream(dev) for _ in range(chains)]
            cs = []
            for c in range(chains):
                chain = ops.SgldChain(arch, SGLD_BATCH, dev, 2e-2, 1e-1, 1e-2, 1e-1, seed=100 + 8 * rank + c)
                net = [(torch.randn(o, i) * (2.0 / (i + o)) ** 0.5, torch.zeros(o)) for i, o in zip(SGLD_DIMS[:-1], SGLD_DIMS[1:])]
                chain.set_state(arch.pack([net])[0], torch.zeros(1))
                chain.bind(X, y)
                cs.append(chain)
            snaps = [torch.empty(arch.stride, device=dev) for _ in range(chains)]

            def sweep(reps, epoch0):
                for r in range(reps):
                    for c in range(chains):
                        with torch.cuda.stream(streams[c]):
                            perm = ops.sgld_perm(100 + c, epoch0 + r, rows, n * SGLD_BATCH, dev)
                            cs[c].run(perm, 0, n)
                            snaps[c].copy_(cs[c].state[: arch.stride])               # thinning snapshot
            torch.cuda.synchronize()
            sweep(2, 0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record()
            for st in streams:
                st.wait_event(e0)
            sweep(reps, 2)
            for st in streams:
                ev = torch.cuda.Event()
                ev.record(st)
                torch.cuda.current_stream().wait_event(ev)
            e1.record()
            torch.cuda.synchronize()
            for c in cs:
                c.check()
            agg = chains * reps * n / (e0.elapsed_time(e1) * 1e-3)
            out[f"{chains}_chains_x_{sms // chains}_ctas"] = {"steps_per_s_total": agg, "steps_per_s_per_chain": agg / chains,
                                                              "state_finite": all(bool(torch.isfinite(c.state).all()) for c in cs)}
            del cs
    finally:
        ops.sgld_set_grid_cap(0)
    del X, y
    torch.cuda.empty_cache()
    return out


def sgld_cfg1_block(dev):
    """cfg 1's training job as the reference runs it (experiments/SGDMC.py: 13-50-1 ReLU, batch 32, 455 training rows, 2500 burn-in
    + 2500 sampling steps, keep_every 50): BNN_SGDMC.train end to end, wall clock, on the single-SM step kernel the library picks
    for a chain this small and -- for comparison -- on the multi-CTA dataflow kernel (BNN_SGLD_SMALL=0)."""
    import torch
    import torch.nn as nn
    from bnn_b200.BNN_SGDMC import BNN_SGDMC
    g = torch.Generator().manual_seed(3)
    X = torch.randn(455, 13, generator=g)
    y = torch.sin(X[:, 0]) + 0.3 * X[:, 1] + 0.1 * torch.randn(455, generator=g)
    out = {"workload": "cfg1 training: BNN_SGDMC.train, 13-50-1 ReLU, batch 32, 455 rows, 2500 + 2500 steps, keep_every 50 "
                       "(launches end at epoch boundaries: <= 15 steps each)"}
    for name, env in (("single_sm", "1"), ("dataflow", "0")):
        os.environ["BNN_SGLD_SMALL"] = env
        try:
            best = None
            for rep in range(3):
                m = BNN_SGDMC(13, nn.ReLU(), [50], nout=1, conf={"device": str(dev), "seed": 5})
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                m.train(X, y)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            rmse, nll = m.validate(X, y, num_samples=50)
            out[name] = {"train_s": best, "steps_per_s": 5000 / best, "us_per_step": best / 5000 * 1e6,
                         "train_rmse": float(rmse), "train_nll": float(nll)}
        finally:
            os.environ.pop("BNN_SGLD_SMALL", None)
    out["kernel"] = ("sgld_small_kernel (one CTA, chain state in shared memory for the whole launch: gather, forward, Gaussian head, "
                     "backward, prior gradient, pSGLD update with in-kernel Philox; CTA barriers only)")
    return out


def dropout_train_block(dev):
    """f2: the fused MC-dropout training kernel at the cfg 3 shape (9-100-100-1 ReLU, batch 32, 4096 rows): steps/s of one launch
    per epoch, next to the reference's recipe on the host cores (oracle port, torch CPU autograd + Adam)."""
    import torch
    from bnn_b200 import ops
    from bnn_b200.layout import Arch
    from oracle import bnn_oracle as O
    dims, B, N, p = [9, 100, 100, 1], 32, 4096, 0.05
    arch = Arch(dims, "relu")
    g = torch.Generator().manual_seed(0)
    X = torch.randn(N, 9, generator=g)
    y = torch.sin(X[:, 0]) + 0.5 * X[:, 1]
    net = [(torch.empty(o, i).uniform_(-1, 1, generator=g) / i ** 0.5, torch.zeros(o)) for i, o in zip(dims[:-1], dims[1:])]
    tr = ops.DropoutTrainer(arch, B, dev, p_drop=p, seed=1)
    tr.set_state(arch.pack([net])[0])
    tr.bind(X.to(dev), y.reshape(-1, 1).to(dev))
    steps = N // B
    perms = [torch.randperm(N, generator=g).to(dev) for _ in range(8)]
    first = float(tr.run(perms[0], 0, steps, want_loss=True).sum())
    tr.run(perms[1], 0, steps)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 24
    e0.record()
    for r in range(reps):
        tr.run(perms[r % 8], 0, steps)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * steps)
    last = float(tr.run(perms[0], 0, steps, want_loss=True).sum())
    n_cpu = 100
    batches = [torch.randperm(N, generator=g)[:B] for _ in range(n_cpu)]
    keeps = [[(torch.rand(B, i) < 1 - p).float() for i in dims[:-1]] for _ in range(n_cpu)]
    cpu, cpu_threads = 0.0, 1
    for nt in (1, os.cpu_count() or 1):
        torch.set_num_threads(nt)
        O.dropout_train_steps(net, X, y, batches[:5], "relu", keeps[:5])
        t0 = time.perf_counter()
        O.dropout_train_steps(net, X, y, batches, "relu", keeps)
        r = n_cpu / (time.perf_counter() - t0)
        if r > cpu:
            cpu, cpu_threads = r, nt
    return {"workload": "cfg3 shape: BNN_Dropout 9-100-100-1 ReLU, batch 32, 4096 rows, launches of one epoch (128 steps)",
            "kernel": "dropout_train_kernel (single-CTA persistent kernel: gather, Philox keep masks, forward, sum-of-squares loss, "
                      "backward, Adam with weight decay; network + moments in shared memory)",
            "steps_per_s": 1e6 / us, "us_per_step": us, "state_finite": bool(torch.isfinite(tr.W).all()),
            "epoch_loss_first": first, "epoch_loss_after": last,
            "cpu": {"steps_per_s": cpu, "cores": cpu_threads, "kind": "port",
                    "sample": f"{n_cpu} steps, torch CPU autograd + Adam, best of 1 / all threads"}}


def config_blocks(dev):
    """BASELINE.json configs 1, 2, 3 and cfg 5's predictive pass through the DROP-IN classes (sample + predict_mv, device
    tensors in and out), each with its own roofline entry.  cfg 1-3 are 7 MFLOP .. 100 GFLOP problems: their time is launch
    latency, which the fractions show."""
    import torch
    import torch.nn as nn
    from bnn_b200 import _lib, ops
    from bnn_b200.BNN_SGDMC import BNN_SGDMC
    from bnn_b200.BNN_BBB import BNN_BBB
    from bnn_b200.BNN_Dropout import BNN_Dropout
    from bnn_b200.BNN_CDropout import BNN_CDropout
    from bnn_b200.bank import SampleBank
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16 = peaks.get("bf16_tflops", 1590.0)
    fp32 = float(_lib.lib().bnn_measure_fp32_tflops())
    out = {}

    def block(name, model, evals, fn, sampling):
        arch = model.arch
        ms = gpu_ms(fn)
        tc = ops.tc_supported(arch, masked=sampling in ("dropout masks", "concrete-dropout masks")) and model.precision in ("auto", "tf32_bf16")
        if model.precision == "auto" and evals * arch.flops_per_eval < ops._AUTO_TC_MIN_FLOP:
            tc = False     # precision="auto" keeps launch-bound problems on the single-launch FP32 kernel
        tf = arch.flops_per_eval * evals / (ms * 1e-3) / 1e12
        peak = bf16 if tc else fp32
        out[name] = {"evals": evals, "ms": ms, "evals_per_s": evals / (ms * 1e-3), "includes": sampling + " + forward + moments",
                     "roofline": {"bound": "tensor" if tc else "fp32", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                                  "kernel": (("fwd_tc_kernel" if len(arch.dims) == 4 and arch.dims[0] <= 64 and max(arch.dims[1:3]) <= 256
                                              else "fwd_tcl_kernel") if tc else "fwd_simt_kernel"),
                                  "flop_per_eval": arch.flops_per_eval}}

    torch.manual_seed(0)
    m = BNN_SGDMC(13, nn.ReLU(), [50], nout=1, conf={"device": str(dev)})
    base = m.arch.pack([m.nn.nn])[0]
    m.nns = SampleBank(m.arch, ((base[None] + 0.05 * torch.randn(100, m.arch.stride)) * m.arch.valid_mask()).to(dev))
    X = torch.randn(51, 13, device=dev)
    block("cfg1 SGDMC 13-50-1 relu, S=100, N=51", m, 100 * 51, lambda: m.predict_mv(X, m.nns), "bank")
    m2 = BNN_BBB(1, nn.Tanh(), [50, 50], conf={"device": str(dev)})
    X2 = torch.linspace(-3, 3, 1000, device=dev).view(-1, 1)
    block("cfg2 BBB 1-50-50-1 tanh, S=1000 draws, N=1000", m2, 1000 * 1000, lambda: m2.predict_mv(X2, m2.sample(1000)), "1000 reparameterised draws")
    X3 = torch.randn(4573, 9, device=dev)
    m3 = BNN_Dropout(9, nn.Tanh(), [100, 100], conf={"device": str(dev)})
    block("cfg3 Dropout 9-100-100-1 tanh, S=1000 masks, N=4573", m3, 1000 * 4573, lambda: m3.predict_mv(X3, m3.sample(1000)), "dropout masks")
    m4 = BNN_CDropout(9, nn.Tanh(), [100, 100], conf={"device": str(dev)})
    block("cfg3 CDropout 9-100-100-1 tanh, S=1000 masks, N=4573", m4, 1000 * 4573, lambda: m4.predict_mv(X3, m4.sample(1000)), "concrete-dropout masks")
    m5 = BNN_SGDMC(90, nn.Tanh(), [512, 512, 512], nout=1, conf={"device": str(dev)})
    base = m5.arch.pack([m5.nn.nn])[0]
    m5.nns = SampleBank(m5.arch, ((base[None] + 0.02 * torch.randn(50, m5.arch.stride)) * m5.arch.valid_mask()).to(dev))
    X5 = torch.randn(100_000, 90, device=dev)
    block("cfg5 predictive 90-512-512-512-1 tanh, S=50, N=100k", m5, 50 * 100_000, lambda: m5.predict_mv(X5, m5.nns), "bank")
    return out


def sgld_cpu_baseline():
    """The reference's sgld_steps body on the host cores: oracle energy (BNN_SGDMC.py:61-74,83) + torch autograd + the oracle's
    pSGLD restatement as the optimizer (pybnn is absent); cfg 5 shape, batch 128."""
    import torch
    from oracle import bnn_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(7)
    Xc, yc = torch.randn(100_000, SGLD_DIMS[0], generator=g), torch.randn(100_000, 1, generator=g)
    params = []
    for i, o in zip(SGLD_DIMS[:-1], SGLD_DIMS[1:]):
        params += [torch.randn(o, i, generator=g) * (2.0 / (i + o)) ** 0.5, torch.zeros(o)]
    params.append(torch.zeros(1))
    Vs = [torch.ones_like(t) for t in params]

    def cpu_step(t):
        idx = torch.randint(0, Xc.shape[0], (SGLD_BATCH,))
        ps = [q.clone().requires_grad_(True) for q in params]
        nn_ = [(ps[2 * l], ps[2 * l + 1]) for l in range(len(SGLD_DIMS) - 1)]
        U = O.sgdmc_neg_log_post(nn_, ps[-1], Xc[idx], yc[idx], "tanh", SGLD_ROWS, 1e-2, 1e-1)
        gs = torch.autograd.grad(U, ps)
        f = float(O.lr_factor(t))
        for k, (q, gq) in enumerate(zip(params, gs)):
            params[k], Vs[k] = O.psgld_step(q, gq, Vs[k], (1e-1 if k == len(params) - 1 else 2e-2) * f, torch.randn_like(q))

    for t in range(5):
        cpu_step(t)
    t0 = time.perf_counter()
    n_cpu = 100
    for t in range(n_cpu):
        cpu_step(5 + t)
    return {"steps_per_s": n_cpu / (time.perf_counter() - t0), "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"{n_cpu} steps, batch {SGLD_BATCH}, torch CPU autograd + oracle pSGLD"}


def sgld_bench(dev, rank, steps):
    """Secondary metric: SGLD steps/s of one pSGLD chain per GPU (cfg 5: synthetic 10M x 90 regression, 3x512 tanh MLP,
    batch 128).  A step = minibatch gather + forward + backward + pSGLD update + LR schedule, all inside the persistent fused
    step kernel (bnn_sgld_run; no library GEMM); the chain runs in launches of keep_every = 50 steps, each followed by the
    thinning snapshot (one bank row copy) and preceded by the epoch-permutation kernel, as BNN_SGDMC.train does."""
    import torch
    import torch.nn as nn
    from bnn_b200 import ops
    from bnn_b200.BNN_SGDMC import BNN_SGDMC
    g = torch.Generator(device=dev).manual_seed(7)
    X = torch.randn(SGLD_ROWS, SGLD_DIMS[0], generator=g, device=dev)
    teacher = BNN_SGDMC(SGLD_DIMS[0], nn.Tanh(), SGLD_DIMS[1:-1], nout=1, conf={"device": str(dev), "seed": 7})
    with torch.no_grad():
        bank = teacher.arch.pack([teacher.nn.nn]).to(dev)
        y = ops.forward(teacher.arch, X[:1_000_000].contiguous(), bank, want_pred=True, want_moments=False, precision="fp32")["pred"].reshape(-1)
        reps = SGLD_ROWS // y.shape[0]
        y = y.repeat(reps) + 0.1 * torch.randn(SGLD_ROWS, generator=g, device=dev)   # teacher on a 1M slice, tiled (synthetic)
    model = BNN_SGDMC(SGLD_DIMS[0], nn.Tanh(), SGLD_DIMS[1:-1], nout=1,
                      conf={"device": str(dev), "seed": 100 + rank, "batch_size": SGLD_BATCH, "keep_every": SGLD_KEEP})
    yv = y.view(-1, 1).contiguous()
    model._chain = model._new_chain(dev)
    model._chain.bind(X, yv)
    model._epoch = 0
    a = model.arch
    n_keep = max(1, steps // SGLD_KEEP)
    snap = torch.empty(n_keep, a.stride, device=dev)

    def run(n_intervals):
        for k in range(n_intervals):
            model.sgld_steps(SGLD_KEEP, SGLD_ROWS)
            snap[k % n_keep].copy_(model._chain.state[: a.stride])       # thinning snapshot (BNN_SGDMC.py:124)

    run(2)                                                               # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(n_keep)
    e1.record()
    torch.cuda.synchronize()
    model._chain.check()
    n_steps = n_keep * SGLD_KEEP
    sps = n_steps / (e0.elapsed_time(e1) * 1e-3)
    loss = float(model._chain.loss)
    out = {"steps_per_s_per_chain": sps, "us_per_step": 1e6 / sps, "steps": n_steps, "batch": SGLD_BATCH, "params": a.n_params + 1,
           "workload": f"cfg5: one pSGLD chain per GPU, synthetic {SGLD_ROWS}x{SGLD_DIMS[0]} regression, "
                       f"{'-'.join(map(str, SGLD_DIMS))} tanh MLP, batch {SGLD_BATCH}, launches of {SGLD_KEEP} steps + thinning snapshot",
           "kernel": "sgld_step_kernel (persistent cooperative kernel, one launch per 50 steps: batch gather, forward, head, backward, "
                     "pSGLD update with in-kernel Philox as a dataflow job list -- tile jobs pulled from a global counter, inputs "
                     "tracked by completion counters, a scheduler warp per CTA, no grid barrier; FP32 FFMA tiles fed by bulk copies; "
                     "no library GEMM, no elementwise launches)",
           "launches_per_50_steps": 3, "last_loss": loss, "state_finite": bool(torch.isfinite(model._chain.state).all()),
           "roofline": {"bound": "latency (dependency chain of 7 tile stages per step, ~10k cycles each; see profiles/r02_sgld_step_notes.md)",
                        "achieved": SGLD_FLOP_PER_STEP * sps / 1e12, "unit": "TFLOP/s", "flop_per_step": SGLD_FLOP_PER_STEP,
                        "update_bytes_per_step": 20 * (a.stride + 1)}}
    del X, y, yv
    torch.cuda.empty_cache()
    return out


_REAL_STDOUT = None


def _quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner goes to fd 1 regardless of
    NCCL_DEBUG_FILE), so fd 1 is pointed at stderr for the whole run and the JSON line is written to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-sgld", action="store_true", help="skip the secondary SGLD steps/s measurement")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config blocks (cfg 1, 2, 3, 5-predictive)")
    ap.add_argument("--no-check", action="store_true", help="skip the spot check against the CPU oracle")
    ap.add_argument("--sgld-steps", type=int, default=2000)
    ap.add_argument("--precision", default="tf32_bf16", choices=["tf32_bf16", "tf32x3", "fp32", "tf32"],
                    help="tf32_bf16: tcgen05 TF32 main term + two BF16 correction terms (FP32-grade, default); tf32x3: 3-pass "
                         "error-compensated TF32 (FP32-grade); fp32: SIMT FFMA; tf32: single pass (looser bound)")
    args = ap.parse_args()
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
