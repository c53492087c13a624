"""Parity of the sm_100a deformable-aggregation kernels against the oracle, torch.grid_sample
(the reference's own fallback path) and the committed reference-op golden."""
import os

import numpy as np
import pytest
import torch

import helpers as h
import oracle
from gaussianformer_b200.ops import DeformableAggregationFunction as DAF
from gaussianformer_b200.synthetic import make_daf_inputs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run(fms, loc, w, backward_seed=None):
    feat, shape, start = DAF.feature_maps_format([f.cuda() for f in fms])
    feat = feat.contiguous().requires_grad_(backward_seed is not None)
    loc_d = loc.cuda().requires_grad_(backward_seed is not None)
    w_d = w.cuda().requires_grad_(backward_seed is not None)
    out = DAF.apply(feat, shape, start, loc_d, w_d)
    grads = None
    if backward_seed is not None:
        g = torch.randn(out.shape, generator=torch.Generator().manual_seed(backward_seed))
        out.backward(g.cuda())
        grads = (g, feat.grad, loc_d.grad, w_d.grad)
    return feat, shape, start, out, grads


@pytest.mark.parametrize("embed,groups,levels", [
    (128, 4, ((12, 20), (6, 10), (3, 5))),        # vectorised path (C % 128 == 0)
    (256, 8, ((9, 7), (5, 4))),                   # two channel slices per warp
    (16, 4, ((7, 12), (4, 6), (2, 3))),           # scalar fallback path
    (24, 3, ((5, 5),)),
])
def test_daf_forward_backward_vs_oracle(embed, groups, levels):
    fms, loc, w = make_daf_inputs(num_anchor=64, num_pts=5, batch=2, num_cams=3, embed_dims=embed,
                                  num_groups=groups, levels=levels, visible_p=0.5, seed=3)
    loc[0, 0, 0] = torch.tensor([0.0, 0.5]); loc[0, 1, 0] = torch.tensor([1.0, 0.5])
    loc[0, 2, 0] = torch.tensor([0.004, 0.996]); loc[0, 3, 1] = torch.tensor([0.999, 0.001])
    loc[1, 4, 2] = torch.tensor([float("nan"), 0.5])
    feat, shape, start, out, (g, gf, gl, gw) = _run(fms, loc, w, backward_seed=11)
    f_np, s_np, st_np = feat.detach().cpu().numpy(), shape.cpu().numpy(), start.cpu().numpy()
    ref = oracle.daf_forward(f_np, s_np, st_np, loc.numpy(), w.numpy(), "f64")
    h.assert_close(out.detach().cpu().numpy(), ref, what="daf out")
    rf, rl, rw = oracle.daf_backward(f_np, s_np, st_np, loc.numpy(), w.numpy(), g.numpy(), "f64")
    h.assert_close(gf.cpu().numpy(), rf, rtol=1e-3, atol=h.grad_tolerance(rf), what="grad feat")
    h.assert_close(gw.cpu().numpy(), rw, rtol=1e-3, atol=h.grad_tolerance(rw), what="grad weights")
    h.assert_close(gl.cpu().numpy(), rl, rtol=1e-3, atol=h.grad_tolerance(rl), what="grad loc")


def test_daf_matches_reference_golden():
    path = os.path.join(GOLD, "ref_daf_small.npz")
    if not os.path.exists(path):
        pytest.skip("golden fixture not generated yet")
    gold = np.load(path)
    levels = ((12, 20), (6, 10), (3, 5))
    fms, loc, w = make_daf_inputs(num_anchor=96, num_pts=5, batch=2, num_cams=3, embed_dims=128, num_groups=4,
                                  levels=levels, visible_p=0.5, seed=4)
    _, _, _, out, (g, gf, gl, gw) = _run(fms, loc, w, backward_seed=int(gold["grad_seed"]))
    h.assert_close(out.detach().cpu().numpy(), gold["out"], what="daf out vs reference op")
    for name, mine, ref in (("feat", gf, gold["grad_feat"]), ("loc", gl, gold["grad_loc"]), ("w", gw, gold["grad_weights"])):
        h.assert_close(mine.cpu().numpy(), ref, rtol=1e-3, atol=h.grad_tolerance(ref), what="grad %s vs reference op" % name)


def test_daf_full_size_properties():
    """BASELINE config-2 shape (P = 25600*9, 6 cams, 4 levels, C = 128): linearity in the weights
    and agreement with a sampled oracle subset (the full oracle would take minutes)."""
    fms, loc, w = make_daf_inputs(num_anchor=25600, num_pts=9, seed=0)
    feat, shape, start, out, _ = _run(fms, loc, w)
    out2 = DAF.apply(feat.detach(), shape, start, loc.cuda(), (2 * w).cuda())
    h.assert_close(out2.cpu().numpy(), 2 * out.detach().cpu().numpy(), what="weight linearity")
    sel = torch.randperm(loc.shape[1], generator=torch.Generator().manual_seed(1))[:4096]
    ref = oracle.daf_forward(feat.detach().cpu().numpy(), shape.cpu().numpy(), start.cpu().numpy(),
                             loc[:, sel].contiguous().numpy(), w[:, sel].contiguous().numpy(), "f64")
    h.assert_close(out.detach()[:, sel.cuda()].cpu().numpy(), ref, what="daf sampled subset")
