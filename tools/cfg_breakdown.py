import sys, time, torch, torch.nn as nn
sys.path.insert(0, '.')
from bnn_b200.BNN_Dropout import BNN_Dropout
from bnn_b200.BNN_BBB import BNN_BBB
from bnn_b200 import ops
def t(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        t0=time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter()-t0)
    return sorted(ts)[len(ts)//2]*1e3
X = torch.randn(4573, 9).cuda()
m = BNN_Dropout(9, nn.Tanh(), [100,100], conf={})
nns = m.sample(1000)
print("cfg3 sample(1000): %.3f ms" % t(lambda: m.sample(1000)))
print("cfg3 predict_mv : %.3f ms" % t(lambda: m.predict_mv(X, nns)))
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
m.predict_mv(X,nns); torch.cuda.synchronize(); e0.record(); m.predict_mv(X,nns); e1.record(); torch.cuda.synchronize(); print("  device time of predict_mv: %.3f ms" % e0.elapsed_time(e1))
m2 = BNN_BBB(1, nn.Tanh(), [50,50], conf={})
X2 = torch.linspace(-3,3,1000).view(-1,1).cuda()
n2 = m2.sample(1000)
print("cfg2 sample(1000): %.3f ms" % t(lambda: m2.sample(1000)))
print("cfg2 predict_mv : %.3f ms" % t(lambda: m2.predict_mv(X2, n2)))
m2.predict_mv(X2,n2); torch.cuda.synchronize(); e0.record(); m2.predict_mv(X2,n2); e1.record(); torch.cuda.synchronize(); print("  device time of predict_mv: %.3f ms" % e0.elapsed_time(e1))
from torch.profiler import profile, ProfilerActivity
for name, fn in (("cfg3 predict_mv", lambda: m.predict_mv(X, nns)), ("cfg2 predict_mv", lambda: m2.predict_mv(X2, n2))):
    fn(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn(); torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    print(name, "kernels:", [(e.name[:40], round(e.device_time, 1)) for e in ev])
