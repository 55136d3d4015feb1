"""Times the fused Bayes-by-Backprop training kernel (bench.py's bbb_train block) on its own."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
print(json.dumps(bench.bbb_train_block(torch.device("cuda:0"))))
