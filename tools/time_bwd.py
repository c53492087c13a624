"""Backward timing: C-ABI call alone vs the autograd path, plus a per-kernel table (torch.profiler)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gaussianformer_b200 import _lib
from gaussianformer_b200.splat import LocalAggregator, LocalAggregatorProb, _make_desc, splat_backward_raw
from gaussianformer_b200.synthetic import make_splat_inputs
dev = "cuda"
def timeit(fn, reps=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for name, cls in (("gs25600_solid", LocalAggregator), ("prob_gs6400", LocalAggregatorProb)):
    kw, inp, _ = make_splat_inputs(name, seed=0, perturb=True)
    m = cls(**kw).to(dev); m.validate = False
    t = {k: v.to(dev) for k, v in inp.items()}
    for k in ("means", "opa", "sem", "cov"): t[k].requires_grad_(True)
    out = m(t["pts"], t["means"], t["opa"], t["sem"], t["scales"], t["cov"])
    outs = list(out) if isinstance(out, (tuple, list)) else [out]
    gs = [torch.randn_like(o) for o in outs]
    wrt = [t["means"], t["opa"], t["sem"], t["cov"]]
    print(name, "autograd bwd ms", timeit(lambda: torch.autograd.grad(outs, wrt, gs, retain_graph=True)))
    # the C-ABI call alone
    cfg = m._cfg()
    G, N, C = t["means"].shape[1], t["pts"].shape[1], t["sem"].shape[2]
    desc = _make_desc(G, N, C, cfg["H"], cfg["W"], cfg["D"], cfg["variant"], cfg["radii_axes"], 9, cfg["pc_min"],
                      cfg["grid_size"], cfg["scale_multiplier"], cfg["radii_min"])
    d = {k: v[0].detach().contiguous() for k, v in t.items()}
    cov = d["cov"].reshape(-1, 9)
    if len(outs) == 1:
        grads_in, saved = (gs[0], None, None), (None, None, None)
    else:
        from gaussianformer_b200.splat import splat_forward_raw
        (lg, bl, de, pr), _ = splat_forward_raw(desc, d["pts"], d["means"], d["opa"], d["sem"], cov, scales=d["scales"])
        grads_in, saved = (gs[0], gs[1], gs[2]), (lg, bl, pr)
    print(name, "C-ABI bwd ms", timeit(lambda: splat_backward_raw(desc, d["pts"], d["means"], d["opa"], d["sem"], cov,
                                                                  grads_in, saved, scales=d["scales"])))
    if "--profile" in sys.argv:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for _ in range(5): torch.autograd.grad(outs, wrt, gs, retain_graph=True)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
