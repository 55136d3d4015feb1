"""Several independent pSGLD chains on ONE GPU (cfg 5: 90-512-512-512-1 tanh, batch 128): each chain's launches go to its own
stream with its CTA count capped (bnn_sgld_set_grid_cap), so the chains' dependency stalls interleave.
    python tools/sgld_multi_chain.py            # 1 x 148, 2 x 74, 3 x 49, 4 x 37 CTAs"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bnn_b200 import ops
from bnn_b200.layout import Arch

dev = torch.device("cuda:0")
dims, act, B, rows, n = [90, 512, 512, 512, 1], "tanh", 128, 1_000_000, 50
arch = Arch(dims, act)
g = torch.Generator(device=dev).manual_seed(0)
X = torch.randn(rows, dims[0], generator=g, device=dev)
y = torch.randn(rows, 1, generator=g, device=dev)
sms = torch.cuda.get_device_properties(0).multi_processor_count
res = []
for chains in (1, 2, 3, 4):
    ops.sgld_set_grid_cap(sms // chains if chains > 1 else 0)
    cs, streams = [], [torch.cuda.Stream(dev) for _ in range(chains)]
    for c in range(chains):
        chain = ops.SgldChain(arch, B, dev, 2e-2, 1e-1, 1e-2, 1e-1, seed=1 + c)
        net = [(torch.randn(o, i) * (2.0 / (i + o)) ** 0.5, torch.zeros(o)) for i, o in zip(dims[:-1], dims[1:])]
        chain.set_state(arch.pack([net])[0], torch.zeros(1))
        chain.bind(X, y)
        cs.append(chain)
    perms = [ops.sgld_perm(1 + c, 0, rows, n * B, dev) for c in range(chains)]
    torch.cuda.synchronize()

    def sweep(reps):
        for r in range(reps):
            for c in range(chains):
                with torch.cuda.stream(streams[c]):
                    cs[c].run(perms[c], 0, n)
    sweep(2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for st in streams:
        st.wait_event(e0)
    sweep(reps)
    for st in streams:
        ev = torch.cuda.Event(); ev.record(st); torch.cuda.current_stream().wait_event(ev)
    e1.record()
    torch.cuda.synchronize()
    for c in cs:
        c.check()
    ms = e0.elapsed_time(e1)
    agg = chains * reps * n / (ms * 1e-3)
    print(f"{chains} chain(s) x {sms // chains} CTAs: {agg:9.0f} steps/s aggregate, {agg / chains:9.0f} per chain, finite {all(bool(torch.isfinite(c.state).all()) for c in cs)}")
    res.append({"chains": chains, "ctas_per_chain": sms // chains, "steps_per_s_aggregate": agg, "steps_per_s_per_chain": agg / chains})
ops.sgld_set_grid_cap(0)
print(json.dumps(res))
