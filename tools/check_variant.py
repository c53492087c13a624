"""Quick hardware check of whatever render variant the environment selects (GF_B200_LIB / GF_B200_ST):
one full-grid sample on exact voxel centres (the fast path of every tile kernel) and one tiny sample (edge tiles, generic
fallbacks) against the fp64 oracle, then a short timing.  Usage: GF_B200_LIB=<variant library> python tools/check_variant.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as h  # noqa: E402

print("variant:", {k: v for k, v in os.environ.items() if k.startswith("GF_B200")}, flush=True)
ok = True
for name, case in (("full grid, voxel centres", ("gs25600_solid", 0, False, dict(G=3000))),
                   ("tiny, perturbed", ("tiny", 1, True, None))):
    kw, inp, variant = h.splat_case(*case)
    m = h.make_module(kw, variant)
    t = h.to_dev(inp)
    out = m(t["pts"], t["means"], t["opa"], t["sem"], t["scales"], t["cov"])
    torch.cuda.synchronize()
    ref = h.oracle_forward(kw, inp, variant)["logits"]
    got = out.cpu().numpy()
    err = np.abs(got - ref)
    bad = int((err > h.ATOL + h.RTOL * np.abs(ref)).sum())
    print(f"{name}: max abs err {err.max():.3e}, outside tolerance {bad} of {err.size}", flush=True)
    ok = ok and bad == 0
    if name.startswith("full"):
        for _ in range(5):
            m(t["pts"], t["means"], t["opa"], t["sem"], t["scales"], t["cov"])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            m(t["pts"], t["means"], t["opa"], t["sem"], t["scales"], t["cov"])
        torch.cuda.synchronize()
        print(f"  module call (G=3000 sample): {(time.perf_counter() - t0) / 50 * 1e3:.4f} ms", flush=True)
print("PARITY OK" if ok else "PARITY FAILED", flush=True)
