"""bench.py's per-config blocks (cfg 1, 2, 3, cfg 5 predictive through the drop-in classes) on their own."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
out = bench.config_blocks(torch.device("cuda:0"))
print(json.dumps({k: {"ms": v["ms"], "evals_per_s": v["evals_per_s"]} for k, v in out.items()}))
