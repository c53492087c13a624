import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers as h
np.set_printoptions(precision=4, suppress=True, linewidth=200)

def run(kw, inp, variant):
    m = h.make_module(kw, variant, validate=False)
    t = h.to_dev(inp)
    out = m(t["pts"], t["means"], t["opa"], t["sem"], t["scales"], t["cov"])
    torch.cuda.synchronize()
    return out.cpu().numpy()

# case 1: all-covering Gaussians with E ~ 1: out[v,c] = sum_g opa_g * sem[g,c]  (constant over voxels)
kw, inp, variant = h.splat_case("tiny", 0, False, dict(G=16, dims=(2, 4, 16), pc_min=(0.0, 0.0, 0.0)))
G = 16
inp["means"][:] = torch.tensor([0.5, 1.0, 4.0])
inp["scales"][:] = 50.0
inp["cov"][:] = torch.eye(3) * 1e-8
inp["opa"][:] = torch.arange(1, G + 1).float()[None] * 1.0
sem = torch.zeros(1, G, 18)
for g in range(G):
    sem[0, g, g % 18] = 1.0
    sem[0, g, 17] = 0.5
inp["sem"] = sem
out = run(kw, inp, variant)
ref = h.oracle_forward(kw, inp, variant)["logits"]
print("case1 ref row0", ref[0])
print("case1 got row0", out[0])
print("case1 got row1", out[1])
print("case1 got row 37", out[37])
print("case1 rows equal to row0:", int((np.abs(out - out[0]) < 1e-3).all(1).sum()), "of", out.shape[0])
print("case1 max abs err", np.abs(out - ref).max())

# case 2: one small Gaussian -> which rows are non-zero
kw, inp, variant = h.splat_case("tiny", 0, False, dict(G=1, dims=(2, 4, 16), pc_min=(0.0, 0.0, 0.0)))
inp["means"][:] = torch.tensor([0.75, 1.25, 5.25])
inp["scales"][:] = 0.2
inp["cov"][:] = torch.eye(3) * 4.0
inp["opa"][:] = 1.0
inp["sem"][:] = torch.arange(1, 19).float()
out = run(kw, inp, variant)
ref = h.oracle_forward(kw, inp, variant)["logits"]
nzr = np.nonzero(np.abs(ref).sum(1))[0]
nzo = np.nonzero(np.abs(out).sum(1))[0]
print("case2 ref nonzero rows", nzr[:40], len(nzr))
print("case2 got nonzero rows", nzo[:40], len(nzo))
if len(nzr):
    print("ref row", nzr[0], ref[nzr[0]][:6])
if len(nzo):
    print("got row", nzo[0], out[nzo[0]][:6])
print("case2 max abs err", np.abs(out - ref).max())
