"""CPU tests of the host-side pieces: layout conversion, synthetic generators, metric."""
import os
import sys

import numpy as np
import pytest
import torch

from gaussianformer_b200.metric import MeanIoU
from gaussianformer_b200.ops.deformable_aggregation import feature_maps_format
from gaussianformer_b200.synthetic import (inverse_covariance, make_daf_inputs, make_splat_inputs, quat_to_rotmat,
                                          voxel_centers)


def test_feature_maps_format_round_trip_and_layout():
    fms = [torch.randn(2, 3, 8, h, w) for h, w in ((6, 10), (3, 5), (2, 2))]
    col, shape, start = feature_maps_format(fms)
    assert col.shape == (2, 3, 60 + 15 + 4, 8)
    assert shape.dtype == torch.int64 and shape.tolist() == [[6, 10], [3, 5], [2, 2]] and start.tolist() == [0, 60, 75]
    # row (start_l + y*w + x) of camera m holds the channel vector of pixel (y, x) of level l
    assert torch.equal(col[1, 2, 60 + 2 * 5 + 3], fms[1][1, 2, :, 2, 3])
    back = feature_maps_format([col, shape, start], inverse=True)
    for a, b in zip(back, fms):
        assert torch.equal(a, b)


def test_voxel_centres_are_x_major_and_in_their_voxel():
    dims, pc_min, gs = (5, 4, 3), (-1.0, 2.0, 0.5), 0.5
    xyz = voxel_centers(dims, pc_min, gs)
    assert xyz.shape == (5, 4, 3, 3)
    flat = xyz.reshape(-1, 3)
    idx = ((flat - torch.tensor(pc_min)) / gs).to(torch.int)
    n = (idx[:, 0] * 4 + idx[:, 1]) * 3 + idx[:, 2]
    assert torch.equal(n, torch.arange(60, dtype=n.dtype))


def test_inverse_covariance_is_spd_and_matches_closed_form():
    g = torch.Generator().manual_seed(0)
    s = 0.1 + torch.rand(50, 3, generator=g)
    q = torch.randn(50, 4, generator=g)
    inv = inverse_covariance(s, q)
    R = quat_to_rotmat(q)
    closed = R.transpose(-1, -2) @ torch.diag_embed(1 / s ** 2) @ R
    assert torch.allclose(inv, closed, rtol=2e-3, atol=1e-3)
    assert (torch.linalg.eigvalsh(closed.double()) > 0).all()


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not mounted")
def test_rotation_matrix_equals_reference_helper():
    sys.path.insert(0, "/root/reference/model/utils")
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_utils", "/root/reference/model/utils/utils.py")
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        sys.path.pop(0)
    q = torch.randn(64, 4, generator=torch.Generator().manual_seed(1))
    assert torch.allclose(quat_to_rotmat(q), ref.get_rotation_matrix(q), atol=1e-6)


def test_synthetic_shapes_and_determinism():
    kw, inp, variant = make_splat_inputs("tiny", seed=3)
    kw2, inp2, _ = make_splat_inputs("tiny", seed=3)
    assert all(torch.equal(inp[k], inp2[k]) for k in inp)
    assert inp["pts"].shape == (1, 50 * 50 * 4, 3) and inp["cov"].shape == (1, 256, 3, 3) and variant == "base"
    fms, loc, w = make_daf_inputs(num_anchor=10, num_pts=3, levels=((4, 4), (2, 2)), embed_dims=8, num_groups=2)
    assert loc.shape == (1, 30, 6, 2) and w.shape == (1, 30, 6, 2, 2)
    s = w.reshape(1, 10, 3 * 6 * 2, 2).sum(2)          # joint softmax over (pts, cams, levels) per group
    assert torch.all((s - 1).abs().lt(1e-5) | s.abs().lt(1e-6))


def test_mean_iou_definition():
    m = MeanIoU([1, 2], empty_label=0)
    pred = torch.tensor([1, 1, 2, 0, 2, 0])
    gt = torch.tensor([1, 2, 2, 0, 0, 1])
    m.after_step(pred, gt)
    miou, iou = m.after_epoch()
    # class 1: correct 1, seen 2, positive 2 -> 1/3 ; class 2: correct 1, seen 2, positive 2 -> 1/3
    assert abs(miou - 100 / 3) < 1e-9
    # occupancy: seen 4 (gt != 0), positive 4, correct 3 -> 3/5
    assert abs(iou - 60.0) < 1e-9


def test_miou_parity_recipe_on_the_oracle():
    """SURVEY.md 8(d) mIoU parity recipe, exercised on the CPU: fp32 vs fp64 oracle logits of the tiny config give
    the same mIoU against the seeded noisy labels up to the voxels whose top two classes tie numerically (each such
    flip moves a class IoU of this 10 000-voxel grid by ~1e-3; at the full 640 000-voxel shape bench.py reports the
    figure); a deliberately damaged prediction does not."""
    import numpy as np
    import helpers as h
    from gaussianformer_b200.metric import miou_parity, synthetic_labels
    kw, inp, variant = h.splat_case("tiny", 0, False)
    ref64 = h.oracle_forward(kw, inp, variant, precision="f64")["logits"]
    ref32 = h.oracle_forward(kw, inp, variant, precision="f32")["logits"]
    C = ref64.shape[1]
    labels, mask = synthetic_labels(ref64, C)
    assert 0.05 < float((labels != torch.as_tensor(ref64).argmax(1)).float().mean()) < 0.15
    r = miou_parity(ref32.argmax(1), ref64.argmax(1), labels, mask, C)
    flips = int((ref32.argmax(1) != ref64.argmax(1)).sum())
    assert flips <= 20 and r["abs_diff"] <= 0.01 * max(flips, 1) and r["ref"][0] > 50.0
    bad = ref64.argmax(1).copy()
    bad[: bad.shape[0] // 4] = 3
    assert miou_parity(bad, ref64.argmax(1), labels, mask, C)["abs_diff"] > 1.0
