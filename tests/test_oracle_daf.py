"""DAF oracle vs torch grid_sample (the reference's own unreachable fallback,
model/encoder/gaussian_encoder/deformable_module.py:307-353) and vs autograd."""
import numpy as np
import torch
import torch.nn.functional as F

import oracle
from gaussianformer_b200.synthetic import make_daf_inputs
from gaussianformer_b200.ops.deformable_aggregation import feature_maps_format


def _dense_daf(feature_maps, loc, w, num_groups):
    """out[b,p,c] = sum_cam gate * sum_lvl w * grid_sample(feat)[c]   (float64, autograd-able)."""
    return oracle.daf_torch_fallback(feature_maps, loc, w, num_groups)


def test_daf_forward_and_backward_match_grid_sample():
    levels = ((7, 12), (4, 6), (2, 3))
    fms, loc, w = make_daf_inputs(num_anchor=40, num_pts=3, batch=2, num_cams=3, embed_dims=16,
                                  num_groups=4, levels=levels, visible_p=0.6, seed=2)
    # put a few samples on the borders / exactly at 0 and 1 (gate is strict)
    loc[0, 0, 0] = torch.tensor([0.0, 0.5]); loc[0, 1, 0] = torch.tensor([1.0, 0.5])
    loc[0, 2, 0] = torch.tensor([0.01, 0.99]); loc[0, 3, 1] = torch.tensor([0.999, 0.001])
    feat, shape, start = feature_maps_format(fms)
    out32 = oracle.daf_forward(feat.numpy(), shape.numpy(), start.numpy(), loc.numpy(), w.numpy(), "f32")
    out64 = oracle.daf_forward(feat.numpy(), shape.numpy(), start.numpy(), loc.numpy(), w.numpy(), "f64")

    fms64 = [f.double().requires_grad_() for f in fms]
    loc64 = loc.double().requires_grad_()
    w64 = w.double().requires_grad_()
    ref = _dense_daf(fms64, loc64, w64, 4)
    # float32 locations are rounded before the oracle multiplies by the level size; use a loose
    # tolerance for the f64 comparison that still catches any indexing slip
    np.testing.assert_allclose(out64, ref.detach().numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(out32, out64, rtol=1e-4, atol=1e-5)

    g = torch.randn(ref.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    ref.backward(g)
    g_feat, g_loc, g_w = oracle.daf_backward(feat.numpy(), shape.numpy(), start.numpy(), loc.numpy(),
                                             w.numpy(), g.float().numpy(), "f64")
    ref_gfeat = feature_maps_format([f.grad for f in fms64])[0].numpy()
    np.testing.assert_allclose(g_feat, ref_gfeat, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(g_w, w64.grad.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(g_loc, loc64.grad.numpy(), rtol=1e-4, atol=2e-4)


def test_fused_oracle_is_the_callers_composition():
    """oracle.daf_fused_* against the reference's literal caller code (deformable_module.py:213-228, :242)
    run through torch autograd around the grid_sample formulation of the op."""
    levels = ((7, 12), (4, 6), (2, 3))
    B, A, K, M, Gr = 2, 12, 3, 3, 4
    fms, loc, _ = make_daf_inputs(num_anchor=A, num_pts=K, batch=B, num_cams=M, embed_dims=16, num_groups=Gr,
                                  levels=levels, visible_p=0.6, seed=7)
    L = len(levels)
    gen = torch.Generator().manual_seed(8)
    logits = torch.randn(B, A, K, M, L, Gr, generator=gen)
    gate = ((loc > 0) & (loc < 1)).all(-1).reshape(B, A, K, M)
    point_mask = gate.clone()
    point_mask[0, 0] = False                                         # an anchor no camera sees: all_miss for every group
    weight_mask = torch.rand(B, A, K, M, L, Gr, generator=gen) > 0.2  # attn_drop mask (:278-279)
    weight_mask[0, 1, :, :, :, 2] = False                             # one all_miss group
    feat, shape, start = feature_maps_format(fms)
    args = (feat.numpy(), shape.numpy(), start.numpy(), loc.numpy())

    for pm, wm in ((point_mask, weight_mask), (point_mask, None), (None, None)):
        fms64 = [f.double().requires_grad_() for f in fms]
        loc64 = loc.double().requires_grad_()
        w = logits.double().requires_grad_()
        # ---- the reference's lines, verbatim modulo names ----
        mask = torch.ones_like(w, dtype=torch.bool)
        if pm is not None:
            mask = pm[..., None, None] & mask
        if wm is not None:
            mask = mask & wm
        all_miss = mask.sum(dim=[2, 3, 4], keepdim=True) == 0
        all_miss = all_miss.expand(-1, -1, K, M, L, -1)
        weights = w.clone()
        weights[~mask] = -torch.inf
        weights[all_miss] = 0.
        weights = weights.flatten(2, 4).softmax(dim=-2).reshape(B, A * K, M, L, Gr)
        weights = weights * (1 - all_miss.flatten(1, 2).float())
        features = _dense_daf(fms64, loc64, weights, Gr).reshape(B, A, K, -1)
        ref = features.sum(dim=2)
        # -------------------------------------------------------
        out = oracle.daf_fused_forward(*args, logits.numpy(), None if pm is None else pm.numpy(),
                                       None if wm is None else wm.numpy(), "f64")
        np.testing.assert_allclose(out, ref.detach().numpy(), rtol=1e-5, atol=1e-5)
        g = torch.randn(ref.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
        ref.backward(g)
        g_feat, g_loc, g_logits = oracle.daf_fused_backward(*args, logits.numpy(), g.float().numpy(),
                                                            None if pm is None else pm.numpy(),
                                                            None if wm is None else wm.numpy(), "f64")
        np.testing.assert_allclose(g_feat, feature_maps_format([f.grad for f in fms64])[0].numpy(), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(g_logits, w.grad.numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(g_loc, loc64.grad.numpy(), rtol=1e-4, atol=2e-4)
        if pm is not None:
            assert np.all(out[0, 0] == 0) and np.all(g_logits[0, 0] == 0)
