"""Shared test helpers: seeded inputs -> oracle results (numpy) in the layouts the ops return."""
import numpy as np
import torch

import oracle
from gaussianformer_b200.synthetic import make_splat_inputs

# fp32 parity gate of the splat op (BASELINE.md "Parity gates"): |new - ref| <= ATOL + RTOL * |ref|
RTOL, ATOL = 1e-4, 1e-5


def splat_case(name, seed=0, perturb=False, overrides=None, variant=None, per_axis=False):
    kw, inp, var = make_splat_inputs(name, seed=seed, perturb=perturb, overrides=overrides)
    variant = variant or var
    if per_axis:
        variant = "prob_fast"
    return kw, inp, variant


def oracle_prep(kw, inp, variant):
    a = {k: v[0].numpy() for k, v in inp.items()}
    dims = (kw["H"], kw["W"], kw["D"])
    pi, mi, radii = oracle.host_prep(a["pts"], a["means"], a["scales"], kw["pc_min"], kw["grid_size"],
                                     kw["scale_multiplier"], radii_min=None if variant == "base" else 1,
                                     per_axis=variant == "prob_fast", dims=None)
    return a, pi, mi, radii, oracle.cov6_from_3x3(a["cov"]), dims


def oracle_forward(kw, inp, variant, precision="f64"):
    a, pi, mi, radii, cov6, dims = oracle_prep(kw, inp, variant)
    if variant == "base":
        out, R = oracle.splat_forward(a["pts"], pi, a["means"], mi, a["opa"], a["sem"], cov6, radii, dims, precision)
        return dict(logits=out, num_pairs=R)
    lg, bl, de, pr, R = oracle.splat_prob_forward(a["pts"], pi, a["means"], mi, a["opa"], a["sem"], cov6, radii, dims,
                                                  precision)
    return dict(logits=lg, bin_logits=bl, density=de, probability=pr, num_pairs=R)


def oracle_backward(kw, inp, variant, grads, saved=None, precision="f64"):
    a, pi, mi, radii, cov6, dims = oracle_prep(kw, inp, variant)
    if variant == "base":
        return oracle.splat_backward(a["pts"], pi, a["means"], mi, a["opa"], a["sem"], cov6, radii, dims, grads[0],
                                     precision)
    return oracle.splat_prob_backward(a["pts"], pi, a["means"], mi, a["opa"], a["sem"], cov6, radii, dims,
                                      saved["logits"], saved["bin_logits"], saved["probability"], grads[0], grads[1],
                                      grads[2], precision)


def make_module(kw, variant, device="cuda", validate=True):
    from gaussianformer_b200.splat import LocalAggregator, LocalAggregatorProb, LocalAggregatorProbFast
    cls = {"base": LocalAggregator, "prob": LocalAggregatorProb, "prob_fast": LocalAggregatorProbFast}[variant]
    m = cls(**kw).to(device)
    m.validate = validate
    return m


def to_dev(inp, device="cuda", requires_grad=False):
    t = {k: v.to(device) for k, v in inp.items()}
    if requires_grad:
        for k in ("means", "opa", "sem", "cov"):
            t[k].requires_grad_(True)
    return t


def assert_close(actual, expected, rtol=RTOL, atol=ATOL, what=""):
    actual = np.asarray(actual, dtype=np.float64)
    expected = np.asarray(expected, dtype=np.float64)
    err = np.abs(actual - expected)
    tol = atol + rtol * np.abs(expected)
    bad = err > tol
    if bad.any():
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.size} outside tolerance; worst at {i}: "
                             f"got {actual[i]!r} want {expected[i]!r} (|err|={err[i]:.3e}, tol={tol[i]:.3e})")


def grad_tolerance(ref):
    """Gradients are sums of up to ~1e5 signed terms: gate on the gradient's own scale."""
    ref = np.asarray(ref, np.float64)
    return 2e-4 * max(1.0, float(np.abs(ref).max()))


def assert_argmax_parity(actual, expected, min_agree=0.9999):
    """Occupancy prediction parity: the arg-max class must agree on >= 99.99 % of the voxels whose
    decision is not a numerical tie (top-two margin of the reference above twice the tolerance);
    a flip on a decided voxel is an error."""
    actual = np.asarray(actual, np.float64)
    expected = np.asarray(expected, np.float64)
    am, ae = actual.argmax(1), expected.argmax(1)
    rows = np.arange(expected.shape[0])
    margin = expected[rows, ae] - expected[rows, am]
    tol = 2 * (ATOL + RTOL * np.abs(expected[rows, ae]))
    decided_flip = (am != ae) & (margin > tol)
    assert not decided_flip.any(), f"{int(decided_flip.sum())} arg-max flips on decided voxels"
    srt = np.sort(expected, axis=1)
    decided = (srt[:, -1] - srt[:, -2]) > 2 * (ATOL + RTOL * np.abs(srt[:, -1]))
    if decided.any():
        assert (am == ae)[decided].mean() >= min_agree
