"""Render-kernel time (events recorded inside the C ABI around the kernel) for a list of configs and batch sizes:
    [GF_B200_LIB=...] python tools/time_render.py gs25600_solid:1 gs25600_solid:4 gs144000:1 gs6400_prob:1
Prints the mean of 30 launches per case (inputs resident, 3 warm-up launches)."""
import ctypes, os, sys
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from gaussianformer_b200.splat import _make_desc, splat_forward_raw  # noqa: E402
from gaussianformer_b200 import _lib  # noqa: E402
from gaussianformer_b200.synthetic import make_splat_inputs  # noqa: E402

L = _lib.lib()
cases = sys.argv[1:] or ["gs25600_solid:1", "gs25600_solid:4", "gs144000:1", "prob_gs6400:1"]
stream = torch.cuda.current_stream()
for case in cases:
    cfg, B = case.split(":"); B = int(B)
    kw, inp, variant = make_splat_inputs(cfg, seed=0, perturb=False)
    t = {k: torch.cat([v] * B, 0).cuda().contiguous() for k, v in inp.items()}
    G, N = t["means"].shape[1], t["pts"].shape[1]
    prob = variant != "base"
    desc = _make_desc(G, N, 18, kw["H"], kw["W"], kw["D"], _lib.GF_SPLAT_PROB if prob else _lib.GF_SPLAT_BASE, 1, 9,
                      kw["pc_min"], kw["grid_size"], float(kw["scale_multiplier"]), 1 if prob else 0, B, 0)
    call = lambda: splat_forward_raw(desc, t["pts"], t["means"], t["opa"], t["sem"], t["cov"].reshape(B, G, 9), scales=t["scales"])
    for _ in range(3):
        call()
    ms = []
    for _ in range(30):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); b.record(stream)
        L.gf_debug_set_render_events(ctypes.c_void_p(a.cuda_event), ctypes.c_void_p(b.cuda_event))
        call()
        L.gf_debug_set_render_events(None, None)
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ms.sort()
    print(f"{cfg} batch {B}: render {1e3 * sum(ms) / len(ms):8.2f} us per launch, {1e3 * sum(ms) / len(ms) / B:8.2f} us per sample "
          f"(min {1e3 * ms[0]:.2f})", flush=True)
