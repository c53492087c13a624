"""Where does the host time of one e2e step go?  (perf_counter around each piece, GPU idle in between)"""
import os, sys, time, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gaussianformer_b200.splat import LocalAggregator
from gaussianformer_b200.synthetic import make_splat_inputs
dev = torch.device("cuda", 0)
kw, inp, _ = make_splat_inputs("gs25600_solid", seed=0)
m = LocalAggregator(**kw).to(dev); m.validate = False
d = {k: v.to(dev) for k, v in inp.items()}
for _ in range(5): m.forward_with_occupancy(d["pts"], d["means"], d["opa"], d["sem"], d["scales"], d["cov"])
torch.cuda.synchronize()
n = 200
t0 = time.perf_counter()
for _ in range(n): m.forward_with_occupancy(d["pts"], d["means"], d["opa"], d["sem"], d["scales"], d["cov"])
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host time per forward_with_occupancy call: %.1f us (queue drained after %.1f us more per call)" % ((t1 - t0) / n * 1e6, (t2 - t1) / n * 1e6))
with torch.no_grad():
    for _ in range(5): m(d["pts"], d["means"], d["opa"], d["sem"], d["scales"], d["cov"]).argmax(dim=1).to(torch.uint8)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): m(d["pts"], d["means"], d["opa"], d["sem"], d["scales"], d["cov"]).argmax(dim=1).to(torch.uint8)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
print("host time per forward + argmax + cast: %.1f us (device time per call %.1f us)" % ((t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6))
host = torch.empty(2_790_434, dtype=torch.float32).pin_memory()
t0 = time.perf_counter()
for _ in range(n): x = host.to(dev, non_blocking=True)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host time per 11 MB H2D launch: %.1f us; total per copy %.1f us" % ((t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6))
