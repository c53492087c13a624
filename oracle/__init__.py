"""ORACLE — TEST INFRASTRUCTURE ONLY.

numpy/ctypes front-end of the CPU restatement in ``oracle/*.c``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package; nothing under ``gaussianformer_b200/`` does, and the product path raises if
its CUDA library is missing instead of falling back to anything here.

Reference anchors (paths relative to the reference tree):

* host preparation of the splat op — ``model/head/localagg/local_aggregate/__init__.py:137-143``
  (prob: ``model/head/localagg_prob/local_aggregate_prob/__init__.py:147-154``, per-axis radii:
  ``model/head/localagg_prob_fast/local_aggregate_prob_fast/__init__.py:151``)
* native semantics — see the headers of ``splat_oracle.c`` and ``daf_oracle.c``.

Parity pin: ``tests/golden/ref_*.npz`` hold outputs of the reference CUDA op itself (built by
``oracle/build_ref.py`` into ``oracle/_ref`` and run on a B200 by ``tests/golden/make_golden_ref.py``);
``tests/test_oracle_golden.py`` checks this oracle against them.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_double, c_float, c_int, c_int32, c_int64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libgf_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement with the committed Makefile (gcc, OpenMP)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.gfo_num_threads.restype = c_int
        for suf in ("_f32", "_f64"):
            getattr(_lib, "gfo_splat_forward" + suf).restype = c_int64
            getattr(_lib, "gfo_splat_prob_forward" + suf).restype = c_int64
    return _lib


def num_threads() -> int:
    return int(lib().gfo_num_threads())


def set_num_threads(n: int) -> int:
    """OpenMP threads of the following oracle calls (torchrun exports OMP_NUM_THREADS=1).  Returns what is in effect."""
    lib().gfo_set_num_threads(int(n))
    return num_threads()


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(c_void_p)


def _real(precision: str):
    if precision == "f32":
        return np.float32, "_f32"
    if precision == "f64":
        return np.float64, "_f64"
    raise ValueError(precision)


# --------------------------------------------------------------------------- host preparation
def host_prep(pts, means, scales, pc_min, grid_size, scale_multiplier, radii_min=None,
              per_axis=False, dims=None):
    """points_int / means_int / radii exactly as the reference's Python wrapper forms them.

    All arithmetic is float32 (numpy weak-scalar rules == torch's tensor-scalar rules), the
    casts truncate toward zero like ``Tensor.to(torch.int)``.  Raises AssertionError where the
    reference asserts (``__init__.py:138,140,142``).
    """
    pts = _f32(pts)
    means = _f32(means)
    scales = _f32(scales)
    pc_min = _f32(pc_min).reshape(1, 3)
    points_int = ((pts - pc_min) / grid_size).astype(np.int32)
    means_int = ((means - pc_min) / grid_size).astype(np.int32)
    if dims is not None:
        H, W, D = dims
        for arr in (points_int, means_int):
            if arr.size:
                assert arr.min() >= 0 and arr[:, 0].max() < H and arr[:, 1].max() < W and arr[:, 2].max() < D
    if per_axis:
        radii = np.ceil(scales * scale_multiplier / grid_size).astype(np.int32)
    else:
        radii = np.ceil(scales.max(axis=-1) * scale_multiplier / grid_size).astype(np.int32)
    if radii_min is not None:
        radii = np.maximum(radii, radii_min).astype(np.int32)
    if radii.size:
        assert radii.min() >= 1
    return points_int, means_int, radii


def cov6_from_3x3(cov):
    """[G,3,3] -> [G,6] taking flat entries [0,4,8,1,5,2] (``__init__.py:143``)."""
    cov = _f32(cov).reshape(-1, 9)
    return np.ascontiguousarray(cov[:, [0, 4, 8, 1, 5, 2]])


def cov6_grad_to_3x3(g6):
    """Scatter a [G,6] gradient back to [G,3,3]: only the six gathered entries receive gradient."""
    g6 = np.asarray(g6)
    out = np.zeros((g6.shape[0], 9), dtype=g6.dtype)
    out[:, [0, 4, 8, 1, 5, 2]] = g6
    return out.reshape(-1, 3, 3)


# --------------------------------------------------------------------------- splat (native level)
def _splat_common(pts, points_int, means, means_int, opa, sem, cov6, radii):
    pts, means, opa, sem, cov6 = map(_f32, (pts, means, opa, sem, cov6))
    points_int, means_int, radii = map(_i32, (points_int, means_int, radii))
    G, N, C = means.shape[0], pts.shape[0], sem.shape[1]
    assert C <= 64
    axes = 3 if radii.ndim == 2 else 1
    return pts, points_int, means, means_int, opa, sem, cov6, radii, G, N, C, axes


def splat_forward(pts, points_int, means, means_int, opa, sem, cov6, radii, dims, precision="f32"):
    """Base forward.  Returns (logits[N,C], num_pairs)."""
    rt, suf = _real(precision)
    pts, points_int, means, means_int, opa, sem, cov6, radii, G, N, C, axes = _splat_common(
        pts, points_int, means, means_int, opa, sem, cov6, radii)
    H, W, D = dims
    out = np.zeros((N, C), dtype=rt)
    R = getattr(lib(), "gfo_splat_forward" + suf)(
        G, N, C, H, W, D, _p(pts), _p(points_int), _p(means), _p(means_int), _p(opa), _p(sem),
        _p(cov6), _p(radii), axes, _p(out))
    return out, int(R)


def splat_backward(pts, points_int, means, means_int, opa, sem, cov6, radii, dims, out_grad,
                   precision="f32"):
    """Base backward.  Returns (g_means[G,3], g_opa[G], g_sem[G,C], g_cov6[G,6])."""
    rt, suf = _real(precision)
    pts, points_int, means, means_int, opa, sem, cov6, radii, G, N, C, axes = _splat_common(
        pts, points_int, means, means_int, opa, sem, cov6, radii)
    H, W, D = dims
    out_grad = _f32(out_grad)
    gm, go = np.zeros((G, 3), rt), np.zeros((G,), rt)
    gs, gc = np.zeros((G, C), rt), np.zeros((G, 6), rt)
    getattr(lib(), "gfo_splat_backward" + suf)(
        G, N, C, H, W, D, _p(pts), _p(points_int), _p(means), _p(means_int), _p(opa), _p(sem),
        _p(cov6), _p(radii), axes, _p(out_grad), _p(gm), _p(go), _p(gs), _p(gc))
    return gm, go, gs, gc


def splat_prob_forward(pts, points_int, means, means_int, opa, sem, cov6, radii, dims,
                       precision="f32"):
    """Prob forward.  Returns (logits[N,C], bin_logits[N], density[N], probability[N], num_pairs)."""
    rt, suf = _real(precision)
    pts, points_int, means, means_int, opa, sem, cov6, radii, G, N, C, axes = _splat_common(
        pts, points_int, means, means_int, opa, sem, cov6, radii)
    H, W, D = dims
    logits = np.zeros((N, C), rt)
    binl, dens, prob = np.zeros((N,), rt), np.zeros((N,), rt), np.zeros((N,), rt)
    R = getattr(lib(), "gfo_splat_prob_forward" + suf)(
        G, N, C, H, W, D, _p(pts), _p(points_int), _p(means), _p(means_int), _p(opa), _p(sem),
        _p(cov6), _p(radii), axes, _p(logits), _p(binl), _p(dens), _p(prob))
    return logits, binl, dens, prob, int(R)


def splat_prob_backward(pts, points_int, means, means_int, opa, sem, cov6, radii, dims,
                        logits, bin_logits, probability, g_logits, g_bin, g_density,
                        precision="f32"):
    rt, suf = _real(precision)
    pts, points_int, means, means_int, opa, sem, cov6, radii, G, N, C, axes = _splat_common(
        pts, points_int, means, means_int, opa, sem, cov6, radii)
    H, W, D = dims
    logits, bin_logits, probability = map(_f32, (logits, bin_logits, probability))
    g_logits, g_bin, g_density = map(_f32, (g_logits, g_bin, g_density))
    gm, go = np.zeros((G, 3), rt), np.zeros((G,), rt)
    gs, gc = np.zeros((G, C), rt), np.zeros((G, 6), rt)
    getattr(lib(), "gfo_splat_prob_backward" + suf)(
        G, N, C, H, W, D, _p(pts), _p(points_int), _p(means), _p(means_int), _p(opa), _p(sem),
        _p(cov6), _p(radii), axes, _p(logits), _p(bin_logits), _p(probability),
        _p(g_logits), _p(g_bin), _p(g_density), _p(gm), _p(go), _p(gs), _p(gc))
    return gm, go, gs, gc


# --------------------------------------------------------------------------- splat (module level)
class LocalAggregatorOracle:
    """numpy mirror of ``LocalAggregator.forward`` for the three variants.

    variant: "base" (``local_aggregate``), "prob" (``local_aggregate_prob``), "prob_fast"
    (``local_aggregate_prob_fast``).  Inputs carry the leading batch dim of 1 like the reference.
    """

    def __init__(self, scale_multiplier, H, W, D, pc_min, grid_size, radii_min=1, variant="base",
                 precision="f32"):
        self.scale_multiplier, self.H, self.W, self.D = scale_multiplier, H, W, D
        self.pc_min = np.asarray(pc_min, np.float32)
        self.grid_size, self.radii_min = grid_size, radii_min
        self.variant, self.precision = variant, precision

    def prep(self, pts, means, scales):
        return host_prep(pts, means, scales, self.pc_min, self.grid_size, self.scale_multiplier,
                         radii_min=None if self.variant == "base" else self.radii_min,
                         per_axis=self.variant == "prob_fast", dims=(self.H, self.W, self.D))

    def forward(self, pts, means, opa, sem, scales, cov):
        assert pts.shape[0] == 1
        pts, means, opa, sem, scales, cov = (np.asarray(a)[0] for a in (pts, means, opa, sem, scales, cov))
        pi, mi, radii = self.prep(pts, means, scales)
        cov6 = cov6_from_3x3(cov)
        dims = (self.H, self.W, self.D)
        if self.variant == "base":
            return splat_forward(pts, pi, means, mi, opa, sem, cov6, radii, dims, self.precision)[0]
        lg, bl, de, _pr, _R = splat_prob_forward(pts, pi, means, mi, opa, sem, cov6, radii, dims,
                                                 self.precision)
        return lg, bl, de


# --------------------------------------------------------------------------- deformable aggregation
def _daf_common(feat, shape, start, loc, weights):
    feat, loc, weights = map(_f32, (feat, loc, weights))
    shape, start = _i32(shape), _i32(start)
    B, M, F, C = feat.shape
    L = shape.shape[0]
    P = loc.shape[1]
    Gr = weights.shape[4]
    assert C % Gr == 0
    return feat, shape, start, loc, weights, (B, M, F, C, L, P, Gr)


def daf_forward(feat, shape, start, loc, weights, precision="f32"):
    rt, suf = _real(precision)
    feat, shape, start, loc, weights, d = _daf_common(feat, shape, start, loc, weights)
    B, M, F, C, L, P, Gr = d
    out = np.zeros((B, P, C), rt)
    getattr(lib(), "gfo_daf_forward" + suf)(B, M, F, C, L, P, Gr, _p(feat), _p(shape), _p(start),
                                            _p(loc), _p(weights), _p(out))
    return out


def daf_backward(feat, shape, start, loc, weights, g_out, precision="f32"):
    rt, suf = _real(precision)
    feat, shape, start, loc, weights, d = _daf_common(feat, shape, start, loc, weights)
    B, M, F, C, L, P, Gr = d
    g_out = _f32(g_out)
    g_feat = np.zeros(feat.shape, rt)
    g_loc = np.zeros(loc.shape, rt)
    g_w = np.zeros(weights.shape, rt)
    getattr(lib(), "gfo_daf_backward" + suf)(B, M, F, C, L, P, Gr, _p(feat), _p(shape), _p(start),
                                             _p(loc), _p(weights), _p(g_out), _p(g_feat), _p(g_loc),
                                             _p(g_w))
    return g_feat, g_loc, g_w


# --------------------------------------------------------------------------- fused caller path of the op
def _fused_weights(logits, point_mask, weight_mask):
    """The weight preparation of DeformableFeatureAggregation.forward
    (model/encoder/gaussian_encoder/deformable_module.py:213-228), float64:
    ``weights[~mask] = -inf; weights[all_miss] = 0; softmax over the flattened (K, M, L) axis;
    weights * (1 - all_miss)``.  ``logits`` is [B, A, K, M, L, Gr]; returns (w [B, A, K, M, L, Gr], mask)."""
    w = np.asarray(logits, np.float64)
    B, A, K, M, L, Gr = w.shape
    mask = np.ones(w.shape, bool)
    if point_mask is not None:
        mask &= np.asarray(point_mask, bool).reshape(B, A, K, M)[..., None, None]     # :213
    if weight_mask is not None:
        mask &= np.asarray(weight_mask, bool).reshape(w.shape)
    all_miss = mask.sum(axis=(2, 3, 4), keepdims=True) == 0                            # :214
    w = np.where(mask, w, -np.inf)                                                     # :216
    w = np.where(all_miss, 0.0, w)                                                     # :217
    flat = w.reshape(B, A, K * M * L, Gr)
    flat = flat - flat.max(axis=2, keepdims=True)
    e = np.exp(flat)
    soft = (e / e.sum(axis=2, keepdims=True)).reshape(B, A, K, M, L, Gr)               # :218
    return soft * (1.0 - all_miss), mask                                               # :228


def daf_fused_forward(feat, shape, start, loc, logits, point_mask=None, weight_mask=None, precision="f64"):
    """deformable_module.py:213-228 + the op (:226) + ``features.sum(dim=2)`` (:242) -> [B, A, C]."""
    B, A, K, M, L, Gr = np.asarray(logits).shape
    w, _ = _fused_weights(logits, point_mask, weight_mask)
    out = daf_forward(feat, shape, start, loc, w.reshape(B, A * K, M, L, Gr), precision)
    return out.reshape(B, A, K, -1).sum(axis=2)


def daf_fused_backward(feat, shape, start, loc, logits, g_out, point_mask=None, weight_mask=None, precision="f64"):
    """Chain rule through :242 (broadcast over key points), the op's own backward and the masked softmax.
    Returns (g_feat, g_loc, g_logits)."""
    B, A, K, M, L, Gr = np.asarray(logits).shape
    w, _ = _fused_weights(logits, point_mask, weight_mask)
    g_pts = np.repeat(np.asarray(g_out, np.float32)[:, :, None, :], K, axis=2).reshape(B, A * K, -1)
    g_feat, g_loc, g_w = daf_backward(feat, shape, start, loc, w.reshape(B, A * K, M, L, Gr), g_pts, precision)
    g_w = np.asarray(g_w, np.float64).reshape(B, A, K, M, L, Gr)
    inner = (w * g_w).sum(axis=(2, 3, 4), keepdims=True)
    # masked entries carry w = 0 (their logit was overwritten, :216-217) and all_miss groups have w = 0 throughout
    return g_feat, g_loc, w * (g_w - inner)


# --------------------------------------------------------------------------- the reference's own PyTorch fallback of the op
def daf_torch_fallback(feature_maps, loc, weights, num_groups):
    """``DeformableFeatureAggregation.feature_sampling`` + ``multi_view_level_fusion``
    (model/encoder/gaussian_encoder/deformable_module.py:307-353), the path the reference takes when its CUDA op is
    not importable, restated on the op's own inputs: ``grid_sample(align_corners=False, padding_mode="zeros")`` per
    level, the strict (0,1) camera gate of the op, weighted sum over cameras and levels.  ``feature_maps``: list of
    ``[B, M, C, h, w]`` tensors, ``loc`` ``[B, P, M, 2]``, ``weights`` ``[B, P, M, L, Gr]``; returns ``[B, P, C]``.
    Differentiable (used as an autograd reference by the tests and timed on the host by bench.py)."""
    import torch
    import torch.nn.functional as F
    B, P, M, _ = loc.shape
    C = feature_maps[0].shape[2]
    gate = ((loc > 0) & (loc < 1)).all(-1)                        # B,P,M
    grid = (loc * 2 - 1).permute(0, 2, 1, 3).reshape(B * M, P, 1, 2)
    out = 0
    for l, fm in enumerate(feature_maps):
        s = F.grid_sample(fm.flatten(0, 1), grid, mode="bilinear", padding_mode="zeros", align_corners=False)
        s = s.reshape(B, M, C, P).permute(0, 3, 1, 2)               # B,P,M,C
        wl = weights[:, :, :, l, :].repeat_interleave(C // num_groups, dim=-1)
        out = out + (s * wl * gate[..., None]).sum(2)
    return out
