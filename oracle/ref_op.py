"""ORACLE — TEST INFRASTRUCTURE ONLY.

Runs the UNMODIFIED reference CUDA ops (built by ``oracle/build_ref.py`` into ``oracle/_ref``) on the
current CUDA device, with the reference's own Python host preparation restated in torch ops on the GPU
(model/head/localagg/local_aggregate/__init__.py:137-143; prob: localagg_prob/local_aggregate_prob/
__init__.py:147-154; per-axis radii: localagg_prob_fast/local_aggregate_prob_fast/__init__.py:151).

Used by ``tests/golden/make_golden_ref.py`` (fixture generation), the full-size op-vs-op parity tests
(``tests/test_parity_full_gpu.py``) and ``bench.py``'s ``ref_cuda_op`` side figure.  Never imported by
``gaussianformer_b200/``.
"""
from __future__ import annotations

import torch

from . import build_ref

_MODULE = {"base": "gf_ref_localagg", "prob": "gf_ref_localagg_prob", "prob_fast": "gf_ref_localagg_prob_fast"}


def available(variant: str) -> bool:
    return build_ref.available(_MODULE[variant])


def host_prep(pts, means, scales, pc_min, grid, mult, radii_min, per_axis):
    points_int = ((pts - pc_min) / grid).to(torch.int)
    means_int = ((means - pc_min) / grid).to(torch.int)
    if per_axis:
        radii = torch.ceil(scales * mult / grid).to(torch.int)
    else:
        radii = torch.ceil(scales.max(dim=-1)[0] * mult / grid).to(torch.int)
    if radii_min is not None:
        radii = radii.clamp(min=radii_min)
    return points_int.contiguous(), means_int.contiguous(), radii.contiguous()


def splat(kw, inp, variant, grads=None, device="cuda"):
    """One sample through the reference op.  ``inp``: dict of CPU tensors with a leading batch dim of 1
    (gaussianformer_b200.synthetic.make_splat_inputs); ``grads``: None (forward only) or the upstream
    gradients ``(g_logits,)`` / ``(g_logits, g_bin, g_density)`` as CPU tensors.

    Returns a dict of CUDA tensors: logits (+ bin_logits, density, probability), num_pairs and, when
    ``grads`` is given, means_grad, opacity_grad, semantics_grad, cov_grad ([G,6] in (xx,yy,zz,xy,yz,xz))."""
    mod = build_ref.load_ref(_MODULE[variant])
    dev = torch.device(device)
    t = {k: v[0].to(dev).contiguous() for k, v in inp.items()}
    pc_min = torch.tensor(kw["pc_min"], dtype=torch.float, device=dev)[None]
    prob = variant != "base"
    pi, mi, radii = host_prep(t["pts"], t["means"], t["scales"], pc_min, kw["grid_size"], kw["scale_multiplier"],
                              1 if prob else None, variant == "prob_fast")
    cov6 = t["cov"].flatten(1)[:, [0, 4, 8, 1, 5, 2]].contiguous()
    H, W, D = kw["H"], kw["W"], kw["D"]
    out = {}
    if not prob:
        R, logits, geom, binning, img = mod.local_aggregate(t["pts"], pi, t["means"], mi, t["opa"], t["sem"], radii,
                                                            cov6, H, W, D)
        out.update(logits=logits, num_pairs=int(R))
        if grads is not None:
            g = grads[0].to(dev).contiguous()
            gm, go, gs, gc = mod.local_aggregate_backward(geom, binning, img, H, W, D, R, t["means"], t["pts"], pi,
                                                          cov6, t["opa"], t["sem"], g)
    else:
        R, logits, binl, dens, probability, geom, binning, img = mod.local_aggregate(
            t["pts"], pi, t["means"], mi, t["opa"], t["sem"], radii, cov6, H, W, D)
        out.update(logits=logits, bin_logits=binl, density=dens, probability=probability, num_pairs=int(R))
        if grads is not None:
            g, gb, gd = (x.to(dev).contiguous() for x in grads)
            gm, go, gs, gc = mod.local_aggregate_backward(geom, binning, img, H, W, D, R, t["means"], t["pts"], pi,
                                                          cov6, t["opa"], t["sem"], logits, binl, dens, probability,
                                                          g, gb, gd)
    if grads is not None:
        out.update(means_grad=gm, opacity_grad=go, semantics_grad=gs, cov_grad=gc)
    torch.cuda.synchronize(dev)
    return out
