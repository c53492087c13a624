"""Per-phase cycle counts of the render kernel (needs a library built with -DGF_RENDER_TIMING):
    python tools/build_variant.py timing -DGF_RENDER_TIMING
    GF_B200_LIB=$PWD/gaussianformer_b200/csrc/variants/libgf_b200_timing.so python tools/phase_timing.py [config] [batch]
Thread 0 of every render CTA adds its clock64() deltas for prologue / Phase A / Phase B / epilogue to four 64-bit
counters in the workspace's status block; this script reads them back and prints the per-CTA averages."""
import ctypes, os, sys
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from gaussianformer_b200.splat import _make_desc, splat_forward_raw  # noqa: E402
from gaussianformer_b200 import _lib  # noqa: E402
from gaussianformer_b200.synthetic import make_splat_inputs  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "gs25600_solid"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
kw, inp, variant = make_splat_inputs(cfg, seed=0, perturb=False)
t = {k: torch.cat([v] * B, 0).cuda().contiguous() for k, v in inp.items()}
G, N = t["means"].shape[1], t["pts"].shape[1]
desc = _make_desc(G, N, 18, kw["H"], kw["W"], kw["D"], _lib.GF_SPLAT_PROB if variant == "prob" else _lib.GF_SPLAT_BASE, 1, 9,
                  kw["pc_min"], kw["grid_size"], float(kw["scale_multiplier"]), 1 if variant == "prob" else 0, B, 0)
for rep in range(3):
    _, ws = splat_forward_raw(desc, t["pts"], t["means"], t["opa"], t["sem"], t["cov"].reshape(B, G, 9), scales=t["scales"])
torch.cuda.synchronize()
c8 = ws[16:80].view(torch.int64).cpu().tolist()
c = c8[:4]
nct = ((kw["H"] + 7) // 8) * ((kw["W"] + 3) // 4) * ((kw["D"] + 15) // 16) * B
names = ("prologue", "phase A", "phase B", "epilogue")
tot = sum(c)
print(f"{cfg} batch {B}: {nct} CTAs; cycles per CTA (thread 0):")
for n, v in zip(names, c):
    print(f"  {n:9s} {v / nct:9.0f}  ({100.0 * v / max(tot, 1):5.1f} %)")
print(f"  total     {tot / nct:9.0f}")
names2 = ("wait records (full)", "arrive + wait empty + issue next", "hit masks (ballots)", "walk (steps)")
tot2 = sum(c8[4:])
print("Phase B of ALL warps, cycles per warp:")
for n, v in zip(names2, c8[4:]):
    print(f"  {n:34s} {v / (4 * nct):9.0f}  ({100.0 * v / max(tot2, 1):5.1f} %)")
