import os, sys, statistics, torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bnn_b200.BNN_Dropout import BNN_Dropout
dev = torch.device("cuda")
def gpu_ms(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)
torch.manual_seed(0)
m = BNN_Dropout(9, nn.Tanh(), [100, 100], conf={"device": "cuda"})
X = torch.randn(4573, 9, device=dev)
def fn():
    return m.predict_mv(X, m.sample(1000))
print(f"cfg3 dropout: {gpu_ms(fn):.4f} ms  PACE={os.environ.get('BNN_TC_PACE')} L2={os.environ.get('BNN_TC_L2PERSIST')}")
