"""Generates tests/golden/ref_*.npz from the REFERENCE CUDA ops themselves.

Run on a B200 box (the reference kernels are CUDA-only):

    gpurun -- 'python tests/golden/make_golden_ref.py gpurun_out/golden'

then copy gpurun_out/golden/*.npz into tests/golden/.  Inputs are NOT stored: they are re-created
bit-identically from (config name, seed) by gaussianformer_b200.synthetic on the CPU.  The host
preparation below restates the reference's Python wrapper with torch ops on the GPU, as the
reference does (model/head/localagg/local_aggregate/__init__.py:137-143).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gaussianformer_b200.synthetic import make_daf_inputs, make_splat_inputs  # noqa: E402
from gaussianformer_b200.ops.deformable_aggregation import feature_maps_format  # noqa: E402
from oracle import build_ref  # noqa: E402

SPLAT_CASES = [  # (fixture, config, seed, overrides, module, per_axis, perturb)
    ("ref_splat_base_tiny", "tiny", 0, None, "gf_ref_localagg", False, False),
    ("ref_splat_base_tiny_perturb", "tiny", 1, None, "gf_ref_localagg", False, True),
    ("ref_splat_prob_tiny", "tiny_prob", 0, None, "gf_ref_localagg_prob", False, False),
    ("ref_splat_probfast_tiny", "tiny_prob", 2, None, "gf_ref_localagg_prob_fast", True, True),
]


def ref_host_prep(pts, means, scales, pc_min, grid, mult, radii_min, per_axis):
    points_int = ((pts - pc_min) / grid).to(torch.int)
    means_int = ((means - pc_min) / grid).to(torch.int)
    if per_axis:
        radii = torch.ceil(scales * mult / grid).to(torch.int)
    else:
        radii = torch.ceil(scales.max(dim=-1)[0] * mult / grid).to(torch.int)
    if radii_min is not None:
        radii = radii.clamp(min=radii_min)
    return points_int.contiguous(), means_int.contiguous(), radii.contiguous()


def run_splat(name, cfg, seed, overrides, modname, per_axis, perturb, outdir):
    kw, inp, variant = make_splat_inputs(cfg, seed=seed, perturb=perturb, overrides=overrides)
    mod = build_ref.load_ref(modname)
    dev = torch.device("cuda")
    t = {k: v[0].to(dev) for k, v in inp.items()}
    pc_min = torch.tensor(kw["pc_min"], dtype=torch.float, device=dev)[None]
    prob = variant == "prob"
    pi, mi, radii = ref_host_prep(t["pts"], t["means"], t["scales"], pc_min, kw["grid_size"],
                                  kw["scale_multiplier"], 1 if prob else None, per_axis)
    cov6 = t["cov"].flatten(1)[:, [0, 4, 8, 1, 5, 2]].contiguous()
    H, W, D = kw["H"], kw["W"], kw["D"]
    gen = torch.Generator().manual_seed(1000 + seed)
    N = t["pts"].shape[0]
    save = {}
    if not prob:
        R, logits, geom, binning, img = mod.local_aggregate(t["pts"], pi, t["means"], mi, t["opa"], t["sem"], radii,
                                                            cov6, H, W, D)
        g = torch.randn(N, 18, generator=gen).to(dev)
        gm, go, gs, gc = mod.local_aggregate_backward(geom, binning, img, H, W, D, R, t["means"], t["pts"], pi, cov6,
                                                      t["opa"], t["sem"], g)
        save.update(logits=logits, num_pairs=np.int64(R))
    else:
        R, logits, binl, dens, probability, geom, binning, img = mod.local_aggregate(
            t["pts"], pi, t["means"], mi, t["opa"], t["sem"], radii, cov6, H, W, D)
        g = torch.randn(N, 18, generator=gen).to(dev)
        gb = torch.randn(N, generator=gen).to(dev)
        gd = torch.randn(N, generator=gen).to(dev)
        gm, go, gs, gc = mod.local_aggregate_backward(geom, binning, img, H, W, D, R, t["means"], t["pts"], pi, cov6,
                                                      t["opa"], t["sem"], logits, binl, dens, probability, g, gb, gd)
        save.update(logits=logits, bin_logits=binl, density=dens, probability=probability, num_pairs=np.int64(R))
    save.update(means_grad=gm, opacity_grad=go, semantics_grad=gs, cov_grad=gc)
    torch.cuda.synchronize()
    arrays = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k, v in save.items()}
    arrays["grad_seed"] = np.int64(1000 + seed)
    np.savez_compressed(os.path.join(outdir, name + ".npz"), **arrays)
    print(name, {k: getattr(v, "shape", v) for k, v in arrays.items()})


def run_daf(outdir):
    mod = build_ref.load_ref("gf_ref_daf")
    dev = torch.device("cuda")
    levels = ((12, 20), (6, 10), (3, 5))
    fms, loc, w = make_daf_inputs(num_anchor=96, num_pts=5, batch=2, num_cams=3, embed_dims=128, num_groups=4,
                                  levels=levels, visible_p=0.5, seed=4)
    feat, shape, start = feature_maps_format(fms)
    feat, loc, w = feat.contiguous().to(dev), loc.to(dev), w.to(dev)
    shape_i, start_i = shape.int().to(dev), start.int().to(dev)
    out = mod.deformable_aggregation_forward(feat, shape_i, start_i, loc, w)
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(2004)).to(dev)
    gf, gl, gw = torch.zeros_like(feat), torch.zeros_like(loc), torch.zeros_like(w)
    mod.deformable_aggregation_backward(feat, shape_i, start_i, loc, w, g, gf, gl, gw)
    torch.cuda.synchronize()
    np.savez_compressed(os.path.join(outdir, "ref_daf_small.npz"), out=out.cpu().numpy(), grad_feat=gf.cpu().numpy(),
                        grad_loc=gl.cpu().numpy(), grad_weights=gw.cpu().numpy(), grad_seed=np.int64(2004))
    print("ref_daf_small", out.shape)


if __name__ == "__main__":
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    for case in SPLAT_CASES:
        run_splat(*case, outdir)
    run_daf(outdir)
