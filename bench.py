#!/usr/bin/env python
"""Headline benchmark: posterior-predictive sample.point MLP evals/s (BASELINE.json metric).

Workload (BASELINE.json configs[3], "synthetic predictive sweep"): N = 1M points, 2x200 ReLU
MLP (D = 16), SGLD samples sharded over the GPUs -- 1250 samples per GPU, so 8 GPUs are the
full S = 10k config and every N runs the same per-GPU shard (weak scaling).  One "step" is one
predict_mv over the whole shard: fused forward + Welford reduction (+ the NCCL all-gather and
Chan merge of the per-point moments when N > 1).

    python bench.py [--gpus N] [--steps K] [--warmup W]           # our CUDA path
    python bench.py --impl reference ...                          # the reference's CPU torch loop

Prints ONE JSON line (see the field notes in DESIGN.md "Measurement").
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
S_PER_GPU = 1250
METRIC = "posterior-predictive sample.point MLP evals/sec"
UNIT = "evals/s"
FLOP_PER_EVAL = 2 * sum(i * o for i, o in zip(DIMS[:-1], DIMS[1:]))   # 86,800


def workload_name(n_gpus):
    return (f"cfg4 synthetic predictive sweep: N={N_POINTS} points x S={S_PER_GPU * n_gpus} SGLD samples "
            f"({S_PER_GPU}/GPU; 8 GPUs = the S=10k config), {DIMS[0]}-{DIMS[1]}-{DIMS[2]}-{DIMS[3]} {ACT} MLP")


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
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
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
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's banner must not share stdout with the JSON line
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    arch = Arch(DIMS, ACT)
    S = S_PER_GPU
    # synthetic inputs: X ~ N(0,1); nets = xavier base + 0.05 N(0,1), one chain shard per rank
    g = torch.Generator(device=dev).manual_seed(1234 + 4)
    X = torch.randn(N_POINTS, DIMS[0], generator=g, device=dev)
    base = arch.pack(cpu_nets(1, 0)[:1])[0].to(dev)
    gw = torch.Generator(device=dev).manual_seed(100 + rank)
    bank = (base[None, :] + 0.05 * torch.randn(S, arch.stride, generator=gw, device=dev)) * arch.valid_mask().to(dev)
    bank = bank.contiguous()
    X_host = X.cpu().pin_memory()
    bank_host = bank.cpu().pin_memory()

    def step_device():
        out = ops.forward(arch, X, bank, want_pred=False, want_moments=(world == 1), want_partials=(world > 1),
                          precision=args.precision)
        if world > 1:
            return merge_shards(out["part_mean"], out["part_m2"], S)
        return out["mean"], out["var"]

    def step_e2e():
        # host-resident inputs: X once, the bank in 5 chunks uploaded behind the evaluation of the previous chunk
        out = ops.forward_streamed(arch, X_host, bank_host, chunk=250, precision=args.precision, want_partials=(world > 1))
        mean, var = merge_shards(out["part_mean"], out["part_m2"], S) if world > 1 else (out["mean"], out["var"])
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
        e0.record()
        for _ in range(steps):
            r = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, r

    clocks = ClockSampler(local) if rank == 0 else None
    ms, (mean, var) = timed(step_device, args.steps, max(args.warmup, 3))
    clk = clocks.stop() if clocks else None
    e2e_steps = max(1, min(args.steps, 3))
    ms_e2e, _ = timed(step_e2e, e2e_steps, 1)

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
        except Exception as e:      # the secondary metric must never take the headline down
            sgld = {"error": repr(e)[:200]}

    evals_per_step = float(S) * n_gpus * N_POINTS
    value = evals_per_step * args.steps / (ms * 1e-3)
    e2e_value = evals_per_step * e2e_steps / (ms_e2e * 1e-3)
    finite = bool(torch.isfinite(mean).all() and torch.isfinite(var).all())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bf16_peak, peak_src = (peaks.get("bf16_tflops_sustained"), "measured (MEASURED_PEAKS.json, sustained)") \
            if peaks.get("bf16_tflops_sustained") else (1400.0, "fallback (B200_PROFILING.md sustained)")
        fp32_peak = float(_lib.lib().bnn_measure_fp32_tflops())
        per_gpu_tflops = FLOP_PER_EVAL * float(S) * N_POINTS * args.steps / (ms * 1e-3) / 1e12
        cpu_val, cpu_t = time_cpu_sample(16, 250_000)
        cores = os.cpu_count() or 1
        tc = args.precision in ("tf32_bf16", "tf32x3", "tf32")
        # tensor work issued per algorithmic MAC, in bf16-equivalent passes: a TF32 MMA pass runs at half the bf16 rate
        # (2 units), a BF16 pass at the full rate (1 unit)
        units = {"tf32_bf16": 2 + 1 + 1, "tf32x3": 3 * 2, "tf32": 2}.get(args.precision, 0)
        scheme = {"tf32_bf16": "one TF32 pass + two BF16 correction passes", "tf32x3": "three TF32 passes",
                  "tf32": "one TF32 pass"}.get(args.precision, "")
        if tc:
            roof = {"bound": "tensor", "achieved": per_gpu_tflops, "peak": bf16_peak, "unit": "TFLOP/s",
                    "frac": per_gpu_tflops / bf16_peak,
                    # dram__bytes_read.sum + dram__bytes_write.sum of this very launch (S=1250 x N=1M) from the committed
                    # ncu --set full capture (profiles/r01_fwd_tc_ncu_full_summary.txt); algorithmic bytes are 0.297 GB -- the
                    # sample-outer loop re-reads X and the float64 per-point state per sample out of L2 (8 % miss to DRAM,
                    # 27 GB/s = 0.3 % of the HBM peak; the kernel is tensor bound)
                    "traffic": 9.888e9 if args.precision == "tf32_bf16" else None, "traffic_unit": "bytes/launch",
                    "kernel": "fwd_tc_kernel<2,relu> (tcgen05 kind::tf32 + kind::f16, cta_group::2, A operand of layer 2 in TMEM)",
                    "flop_per_eval": FLOP_PER_EVAL,
                    "peak_source": peak_src + "; dense bf16",
                    "note": f"achieved = ALGORITHMIC flop/s (2*sum in*out per eval).  {args.precision} issues {scheme} per algorithmic "
                            f"MAC = {units} bf16-equivalent units, so frac is capped at {1.0 / units:.3f}; tensor_pipe_frac below is "
                            "the share of the bf16 peak the issued MMAs occupy",
                    "tensor_pipe_frac": units * per_gpu_tflops / bf16_peak}
        else:
            roof = {"bound": "fp32", "achieved": per_gpu_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
                    "frac": per_gpu_tflops / fp32_peak if fp32_peak > 0 else None, "traffic": None,
                    "kernel": "fwd_simt_kernel<8>", "flop_per_eval": FLOP_PER_EVAL,
                    "peak_source": "FP32 FFMA microbenchmark measured live (bnn_measure_fp32_tflops); "
                                   "MEASURED_PEAKS.json has no FP32-SIMT figure",
                    "frac_of_bf16_tensor_peak": per_gpu_tflops / bf16_peak, "bf16_peak": bf16_peak}
        launches_per_step = (3 if tc else 2) + (1 if world > 1 else 0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"tf32_bf16": "tf32+bf16 (TF32 main term + two BF16 correction terms, FP32 accumulate; parity 1e-5)",
                                           "tf32x3": "tf32x3 (3-pass error-compensated TF32, FP32 accumulate; parity 1e-5)",
                                           "tf32": "tf32", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": {"workload": workload_name(n_gpus), "precision": args.precision,
                       "l2": f"inputs {(bank.numel() + X.numel()) * 4 / 1e6:.0f} MB per GPU > 126 MB L2 (no flush needed)",
                       "parallelism": f"samples sharded x{n_gpus}, all-gather of float64 (mean, M2) + Chan merge"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int((bank.numel() + X.numel()) * 4),
                    "d2h_bytes_per_step": int(2 * N_POINTS * DIMS[-1] * 4), "ms_per_step": ms_e2e / e2e_steps,
                    "steps": e2e_steps, "api": "bnn_b200.ops.forward_streamed on pinned host tensors (H2D of X + bank in 5 chunks overlapped with the evaluation, D2H of mean/var, per step)"},
            "gpu_launches": args.steps * launches_per_step,
            "roofline": roof,
            "fp32_ffma_peak_tflops": fp32_peak,
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"S=16 samples x N=250000 points of the same workload, torch CPU loop, "
                                       f"{cores} threads, {cpu_t:.1f} s"},
            "clocks": clk, "outputs_finite": finite, "sgld": sgld,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


SGLD_DIMS, SGLD_BATCH, SGLD_ROWS = [90, 512, 512, 512, 1], 128, 10_000_000     # BASELINE.json configs[4]


def sgld_bench(dev, rank, steps, cpu=True):
    """Secondary metric: SGLD steps/s of one pSGLD chain per GPU (cfg 5: synthetic 10M x 90 regression, 3x512 tanh MLP,
    batch 128).  A step = minibatch gather + forward + backward (torch autograd captured in a CUDA graph: the stop-gap
    SURVEY.md 8f ranks next) + the fused pSGLD update kernel + LR schedule."""
    import torch
    import torch.nn as nn
    from bnn_b200.BNN_SGDMC import BNN_SGDMC
    g = torch.Generator(device=dev).manual_seed(7)
    X = torch.randn(SGLD_ROWS, SGLD_DIMS[0], generator=g, device=dev)
    teacher = BNN_SGDMC(SGLD_DIMS[0], nn.Tanh(), SGLD_DIMS[1:-1], nout=1, conf={"device": str(dev), "seed": 7})
    with torch.no_grad():
        bank = teacher.arch.pack([teacher.nn.nn]).to(dev)
        from bnn_b200 import ops
        y = ops.forward(teacher.arch, X[:1_000_000].contiguous(), bank, want_pred=True, want_moments=False)["pred"].reshape(-1)
        reps = SGLD_ROWS // y.shape[0]
        y = y.repeat(reps) + 0.1 * torch.randn(SGLD_ROWS, generator=g, device=dev)   # teacher on a 1M slice, tiled (synthetic)
    model = BNN_SGDMC(SGLD_DIMS[0], nn.Tanh(), SGLD_DIMS[1:-1], nout=1,
                      conf={"device": str(dev), "seed": 100 + rank, "batch_size": SGLD_BATCH})
    model._chain_init(dev)
    model._perm, model._perm_pos = None, 0
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    yv = y.view(-1, 1)
    model.sgld_steps(50, X, yv, gen)                    # warm-up (builds the graph)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model.sgld_steps(steps, X, yv, gen)
    e1.record()
    torch.cuda.synchronize()
    sps = steps / (e0.elapsed_time(e1) * 1e-3)
    out = {"steps_per_s_per_chain": sps, "steps": steps, "batch": SGLD_BATCH, "params": model.arch.n_params + 1,
           "workload": f"cfg5: one pSGLD chain per GPU, synthetic {SGLD_ROWS}x{SGLD_DIMS[0]} regression, "
                       f"{'-'.join(map(str, SGLD_DIMS))} tanh MLP, batch {SGLD_BATCH}",
           "gradient": "hand-derived on the packed vector: bnn_sgld_gather + 10 cuBLAS SGEMMs + 6 activation kernels + bnn_sgld_prior + bnn_sgld_head, whole step = one CUDA-graph replay of 21 launches", "update": "bnn_psgld_step (fused, Philox in-kernel)",
           "update_bytes_per_step": 20 * (model.arch.stride + 1)}
    if cpu and rank == 0:
        # CPU: the reference's sgld_steps body with the oracle's pSGLD restatement as the optimizer (pybnn is absent)
        from oracle import bnn_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        Xc, yc = X[:100_000].cpu(), yv[:100_000].cpu()
        net = [(W.clone(), b.clone()) for W, b in teacher.arch.unpack(bank[0].cpu())]
        params = [t for pair in net for t in pair] + [torch.zeros(1)]
        Vs = [torch.ones_like(t) for t in params]
        def cpu_step(t):
            idx = torch.randint(0, Xc.shape[0], (SGLD_BATCH,))
            ps = [q.clone().requires_grad_(True) for q in params]
            nn_ = [(ps[0], ps[1]), (ps[2], ps[3]), (ps[4], ps[5]), (ps[6], ps[7])]
            U = O.sgdmc_neg_log_post(nn_, ps[8], Xc[idx], yc[idx], "tanh", SGLD_ROWS, 1e-2, 1e-1)
            gs = torch.autograd.grad(U, ps)
            f = float(O.lr_factor(t))
            for k, (q, gq) in enumerate(zip(params, gs)):
                params[k], Vs[k] = O.psgld_step(q, gq, Vs[k], (1e-1 if k == 8 else 2e-2) * f, torch.randn_like(q))
        for t in range(5):
            cpu_step(t)
        t0 = time.perf_counter()
        n_cpu = 100
        for t in range(n_cpu):
            cpu_step(5 + t)
        out["cpu_steps_per_s"] = n_cpu / (time.perf_counter() - t0)
        out["cpu_cores"] = os.cpu_count() or 1
    del X, y
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
