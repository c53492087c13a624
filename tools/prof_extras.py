"""Runs each side op a few times (for `ncu --metrics gpu__time_duration.sum` launch lists)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gaussianformer_b200.splat import LocalAggregator, LocalAggregatorProb
from gaussianformer_b200.synthetic import make_daf_inputs, make_splat_inputs
from gaussianformer_b200.ops import DeformableAggregationFunction as DAF
dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = 3
if which in ("all", "bwd"):
    kw, inp, _ = make_splat_inputs("gs25600_solid", seed=0, perturb=True)
    m = LocalAggregator(**kw).to(dev); m.validate = False
    t = {k: v.to(dev) for k, v in inp.items()}
    for k in ("means", "opa", "sem", "cov"): t[k].requires_grad_(True)
    out = m(t["pts"], t["means"], t["opa"], t["sem"], t["scales"], t["cov"])
    g = torch.randn_like(out)
    for _ in range(reps): torch.autograd.grad(out, [t["means"], t["opa"], t["sem"], t["cov"]], g, retain_graph=True)
if which in ("all", "prob"):
    kw, inp, _ = make_splat_inputs("prob_gs6400", seed=0, perturb=True)
    m = LocalAggregatorProb(**kw).to(dev); m.validate = False
    t = {k: v.to(dev) for k, v in inp.items()}
    for k in ("means", "opa", "sem", "cov"): t[k].requires_grad_(True)
    for _ in range(reps):
        lg, bl, de = m(t["pts"], t["means"], t["opa"], t["sem"], t["scales"], t["cov"])
        torch.autograd.grad([lg.sum() + bl.sum() + de.sum()], [t["means"], t["opa"], t["sem"], t["cov"]])
if which in ("all", "daf"):
    fms, loc, w = make_daf_inputs(seed=0)
    feat, shape, start = DAF.feature_maps_format([f.to(dev) for f in fms])
    feat = feat.contiguous().requires_grad_(); loc = loc.to(dev).requires_grad_(); w = w.to(dev).requires_grad_()
    for _ in range(reps):
        o = DAF.apply(feat, shape, start, loc, w)
        o.sum().backward()
torch.cuda.synchronize()
