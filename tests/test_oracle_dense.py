"""The C oracle against independent dense float64 formulas + autograd on tiny problems."""
import numpy as np
import pytest
import torch

import oracle
from gaussianformer_b200.synthetic import make_splat_inputs
import dense_ref


def _prep(name, seed, variant=None, G=None, perturb=False, per_axis=False):
    kw, inp, var = make_splat_inputs(name, seed=seed, perturb=perturb,
                                     overrides=dict(G=G) if G else None)
    variant = variant or var
    dims = (kw["H"], kw["W"], kw["D"])
    a = {k: v[0].numpy() for k, v in inp.items()}
    pi, mi, radii = oracle.host_prep(a["pts"], a["means"], a["scales"], kw["pc_min"], kw["grid_size"],
                                     kw["scale_multiplier"], radii_min=None if variant == "base" else 1,
                                     per_axis=per_axis, dims=dims)
    cov6 = oracle.cov6_from_3x3(a["cov"])
    return a, pi, mi, radii, cov6, dims


def _t(x, dt=torch.float64):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dt)


@pytest.mark.parametrize("per_axis", [False, True])
@pytest.mark.parametrize("perturb", [False, True])
def test_base_forward_matches_dense(per_axis, perturb):
    a, pi, mi, radii, cov6, dims = _prep("tiny", 3, G=96, perturb=perturb, per_axis=per_axis)
    out64, R = oracle.splat_forward(a["pts"], pi, a["means"], mi, a["opa"], a["sem"], cov6, radii, dims, "f64")
    out32, R32 = oracle.splat_forward(a["pts"], pi, a["means"], mi, a["opa"], a["sem"], cov6, radii, dims, "f32")
    ref = dense_ref.base_forward(_t(a["pts"]), _t(pi, torch.long), _t(a["means"]), _t(mi, torch.long),
                                 _t(a["opa"]), _t(a["sem"]), _t(cov6), _t(radii, torch.long)).numpy()
    mask = dense_ref.inclusion_mask(_t(pi, torch.long), _t(mi, torch.long), _t(radii, torch.long))
    assert R == R32 == int(mask.sum())
    np.testing.assert_allclose(out64, ref, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(out32, ref, rtol=1e-4, atol=1e-5)


def test_base_backward_matches_autograd():
    a, pi, mi, radii, cov6, dims = _prep("tiny", 5, G=64)
    rng = np.random.default_rng(0)
    gout = rng.standard_normal((a["pts"].shape[0], 18)).astype(np.float32)
    gm, go, gs, gc = oracle.splat_backward(a["pts"], pi, a["means"], mi, a["opa"], a["sem"], cov6, radii,
                                           dims, gout, "f64")
    means = _t(a["means"]).requires_grad_()
    opa = _t(a["opa"]).requires_grad_()
    sem = _t(a["sem"]).requires_grad_()
    c6 = _t(cov6).requires_grad_()
    out = dense_ref.base_forward(_t(a["pts"]), _t(pi, torch.long), means, _t(mi, torch.long), opa, sem, c6,
                                 _t(radii, torch.long))
    out.backward(_t(gout))
    for mine, ref in ((gm, means.grad), (go, opa.grad), (gs, sem.grad), (gc, c6.grad)):
        np.testing.assert_allclose(mine, ref.numpy(), rtol=1e-9, atol=1e-11)
    gm32, go32, gs32, gc32 = oracle.splat_backward(a["pts"], pi, a["means"], mi, a["opa"], a["sem"], cov6,
                                                   radii, dims, gout, "f32")
    for m32, m64 in ((gm32, gm), (go32, go), (gs32, gs), (gc32, gc)):
        np.testing.assert_allclose(m32, m64, rtol=2e-3, atol=2e-4 * np.abs(m64).max())


@pytest.mark.parametrize("per_axis", [False, True])
def test_prob_forward_matches_dense(per_axis):
    a, pi, mi, radii, cov6, dims = _prep("tiny_prob", 7, per_axis=per_axis)
    lg, bl, de, pr, R = oracle.splat_prob_forward(a["pts"], pi, a["means"], mi, a["opa"], a["sem"], cov6,
                                                  radii, dims, "f64")
    ref = dense_ref.prob_forward(_t(a["pts"]), _t(pi, torch.long), _t(a["means"]), _t(mi, torch.long),
                                 _t(a["opa"]), _t(a["sem"]), _t(cov6), _t(radii, torch.long))
    # voxels whose Z sits within rounding of the 1e-9 switch may legitimately take either branch
    z = ref[3].numpy()
    stable = np.abs(z - 1e-9) > 1e-12
    np.testing.assert_allclose(lg[stable], ref[0].numpy()[stable], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(bl, ref[1].numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(de, ref[2].numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(pr, z, rtol=1e-10, atol=1e-20)
    # fallback branch: uniform over the first 17 classes, last channel zero
    empty = z <= 1e-9
    if empty.any():
        assert np.allclose(lg[empty & stable][:, :17], 1.0 / 17) and np.all(lg[empty & stable][:, 17] == 0)


def test_prob_backward_matches_autograd_away_from_switches():
    """The reference's prob backward is autograd-exact except for the +1e-9 guard in the
    bin term and the skipped logits branch; with E well below 1 and Z > 1e-9 everywhere that a
    gradient flows, it must agree with autograd of the dense formula."""
    a, pi, mi, radii, cov6, dims = _prep("tiny_prob", 11)
    N = a["pts"].shape[0]
    lg, bl, de, pr, _ = oracle.splat_prob_forward(a["pts"], pi, a["means"], mi, a["opa"], a["sem"], cov6,
                                                  radii, dims, "f64")
    rng = np.random.default_rng(1)
    g_lg = rng.standard_normal((N, 18)).astype(np.float32)
    g_bl = rng.standard_normal(N).astype(np.float32)
    g_de = rng.standard_normal(N).astype(np.float32)
    g_lg[pr <= 1e-6] = 0   # keep away from the Z switch
    # the oracle consumes float32 saved outputs like the reference; feed it float64-accurate ones
    means = _t(a["means"]).requires_grad_()
    opa = _t(a["opa"]).requires_grad_()
    sem = _t(a["sem"]).requires_grad_()
    c6 = _t(cov6).requires_grad_()
    out = dense_ref.prob_forward(_t(a["pts"]), _t(pi, torch.long), means, _t(mi, torch.long), opa, sem, c6,
                                 _t(radii, torch.long))
    (out[0] * _t(g_lg)).sum().add((out[1] * _t(g_bl)).sum()).add((out[2] * _t(g_de)).sum()).backward()
    gm, go, gs, gc = oracle.splat_prob_backward(a["pts"], pi, a["means"], mi, a["opa"], a["sem"], cov6, radii,
                                                dims, lg, bl, pr, g_lg, g_bl, g_de, "f64")
    # saved outputs were rounded to float32 on the way in => ~1e-7 relative agreement
    for mine, ref in ((gm, means.grad), (go, opa.grad), (gs, sem.grad), (gc, c6.grad)):
        ref = ref.numpy()
        np.testing.assert_allclose(mine, ref, rtol=5e-5, atol=5e-6 * max(1.0, np.abs(ref).max()))
