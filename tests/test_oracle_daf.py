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
    B, P, M, _ = loc.shape
    C = feature_maps[0].shape[2]
    gate = ((loc > 0) & (loc < 1)).all(-1)                        # B,P,M
    grid = (loc * 2 - 1).permute(0, 2, 1, 3).reshape(B * M, P, 1, 2)
    out = 0
    for l, fm in enumerate(feature_maps):
        s = F.grid_sample(fm.flatten(0, 1), grid, mode="bilinear", padding_mode="zeros",
                          align_corners=False)                      # B*M,C,P,1
        s = s.reshape(B, M, C, P).permute(0, 3, 1, 2)               # B,P,M,C
        wl = w[:, :, :, l, :].repeat_interleave(C // num_groups, dim=-1)   # B,P,M,C
        out = out + (s * wl * gate[..., None]).sum(2)
    return out


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
