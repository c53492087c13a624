"""Times the side ops with CUDA events (bwd, prob fwd/bwd, DAF fwd/bwd) and the e2e pieces."""
import os, sys, time, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gaussianformer_b200.splat import LocalAggregator, LocalAggregatorProb
from gaussianformer_b200.synthetic import make_daf_inputs, make_splat_inputs
from gaussianformer_b200.ops import DeformableAggregationFunction as DAF
dev = "cuda"
def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
kw, inp, _ = make_splat_inputs("gs25600_solid", seed=0, perturb=True)
m = LocalAggregator(**kw).to(dev); m.validate = False
t = {k: v.to(dev) for k, v in inp.items()}
for k in ("means", "opa", "sem", "cov"): t[k].requires_grad_(True)
out = m(t["pts"], t["means"], t["opa"], t["sem"], t["scales"], t["cov"])
g = torch.randn_like(out)
print("base bwd ms", timeit(lambda: torch.autograd.grad(out, [t["means"], t["opa"], t["sem"], t["cov"]], g, retain_graph=True)))
kwp, inpp, _ = make_splat_inputs("prob_gs6400", seed=0, perturb=True)
mp = LocalAggregatorProb(**kwp).to(dev); mp.validate = False
tp = {k: v.to(dev) for k, v in inpp.items()}
for k in ("means", "opa", "sem", "cov"): tp[k].requires_grad_(True)
print("prob fwd ms", timeit(lambda: mp(tp["pts"], tp["means"], tp["opa"], tp["sem"], tp["scales"], tp["cov"])))
lg, bl, de = mp(tp["pts"], tp["means"], tp["opa"], tp["sem"], tp["scales"], tp["cov"])
gl, gb, gd = torch.randn_like(lg), torch.randn_like(bl), torch.randn_like(de)
print("prob bwd ms", timeit(lambda: torch.autograd.grad([lg, bl, de], [tp["means"], tp["opa"], tp["sem"], tp["cov"]], [gl, gb, gd], retain_graph=True), reps=5))
fms, loc, w = make_daf_inputs(seed=0)
feat, shape, start = DAF.feature_maps_format([f.to(dev) for f in fms])
feat = feat.contiguous().requires_grad_(); loc = loc.to(dev).requires_grad_(); w = w.to(dev).requires_grad_()
dev_maps = [f.to(dev) for f in fms]
print("format fused ms", timeit(lambda: DAF.feature_maps_format(dev_maps)[0]))
from gaussianformer_b200.ops.deformable_aggregation import _FeatureMapsToTable
print("format kernel via autograd.Function only ms", timeit(lambda: _FeatureMapsToTable.apply(*dev_maps)))
def torch_format():
    bs, cams, ch = dev_maps[0].shape[:3]
    return torch.cat([m.reshape(bs, cams, ch, -1) for m in dev_maps], dim=-1).permute(0, 1, 3, 2).contiguous()
print("format reference route (cat + permute + contiguous) ms", timeit(torch_format))
print("daf fwd ms", timeit(lambda: DAF.apply(feat, shape, start, loc, w)))
o = DAF.apply(feat, shape, start, loc, w); go = torch.randn_like(o)
print("daf bwd ms (incl. 3 zero fills)", timeit(lambda: torch.autograd.grad(o, [feat, loc, w], go, retain_graph=True), reps=5))
# e2e pieces
host = {k: v.pin_memory() for k, v in inp.items()}
def h2d():
    return {k: v.to(dev, non_blocking=True) for k, v in host.items()}
print("h2d ms", timeit(h2d))
d = h2d()
print("fwd+occ ms", timeit(lambda: m.forward_with_occupancy(d["pts"], d["means"], d["opa"], d["sem"], d["scales"], d["cov"])))
occ = torch.empty(640000, dtype=torch.uint8, device=dev); ph = torch.empty(640000, dtype=torch.uint8).pin_memory()
print("d2h ms", timeit(lambda: ph.copy_(occ, non_blocking=True)))
