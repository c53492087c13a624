"""Host-side mirror of the reference's three splat packages.

``LocalAggregator`` keeps the reference's constructor kwargs, buffer name (``pc_min``) and call
signature (``model/head/localagg/local_aggregate/__init__.py:108-161``,
``model/head/localagg_prob/local_aggregate_prob/__init__.py:118-169``,
``model/head/localagg_prob_fast/local_aggregate_prob_fast/__init__.py:151``) so that
``GaussianHead`` (``model/head/gaussian_head.py:30-39,157-163``) can construct and call it
unchanged.  The arithmetic — including the reference's Python-side preparation (voxel indices,
radii, 3x3 -> 6 gather) — runs in the sm_100a kernels behind the C ABI (``include/gf_b200.h``).

Differences from the reference, all supersets:

* batch sizes > 1 are accepted (the reference asserts ``B == 1``); ``B == 1`` returns the same
  squeezed shapes as the reference;
* no host synchronisation inside the op except one status-word read when ``validate=True``
  (the reference performs >= 7 ``.min()/.max()`` syncs plus a blocking memcpy);
* inputs in any float dtype are computed in fp32 (the reference would throw on half tensors).
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.nn as nn

from . import _lib
from ._lib import SplatDesc, SplatGrads, SplatInputs, SplatOutputs

_COV_IDX = (0, 4, 8, 1, 5, 2)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _f32c(t):
    t = t.detach()
    if t.dtype is torch.float32 and t.is_contiguous():
        return t                      # the common case costs two attribute reads
    return t.contiguous().float()


_DESC_CACHE = {}
_WS_BYTES = {}   # id(cached desc) -> forward workspace bytes


def _cached_desc(key, *args):
    """ctypes descriptors are immutable after construction here; build each distinct one once."""
    d = _DESC_CACHE.get(key)
    if d is None:
        d = _make_desc(*args)
        if len(_DESC_CACHE) < 256:
            _DESC_CACHE[key] = d
    return d


def _make_desc(G, N, C, H, W, D, variant, radii_axes, cov_stride, pc_min, grid_size, scale_multiplier, radii_min):
    d = SplatDesc()
    d.G, d.N, d.C, d.H, d.W, d.D = G, N, C, H, W, D
    d.variant, d.radii_axes, d.cov_stride = variant, radii_axes, cov_stride
    d.pc_min[0], d.pc_min[1], d.pc_min[2] = pc_min
    d.grid_size, d.scale_multiplier, d.radii_min = grid_size, scale_multiplier, radii_min
    return d


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "gaussianformer_b200 ops are CUDA-only (sm_100a); got a CPU tensor. There is no CPU fallback.")


def splat_forward_raw(desc, pts, means, opa, sem, cov, *, points_int=None, means_int=None, radii=None, scales=None,
                      argmax_out=None):
    """One sample through ``gf_splat_forward``.  Returns (outputs tuple, workspace tensor).
    ``argmax_out``: optional uint8 ``[N]`` tensor that receives the fused arg-max class."""
    L = _lib.lib()
    dev = pts.device
    with torch.cuda.device(dev):
        N, C = desc.N, desc.C
        logits = torch.empty((N, C), dtype=torch.float32, device=dev)
        prob = desc.variant == _lib.GF_SPLAT_PROB
        if prob:
            aux = torch.empty((3, N), dtype=torch.float32, device=dev)
            binl, dens, probability = aux[0], aux[1], aux[2]
        else:
            binl = dens = probability = None
        ws_bytes = _WS_BYTES.get(id(desc))
        if ws_bytes is None:
            ws_bytes = L.gf_splat_forward_workspace_bytes(ctypes.byref(desc))
            if ws_bytes == 0:
                raise _lib.GfError(L.gf_last_error().decode())
            if id(desc) in map(id, _DESC_CACHE.values()):
                _WS_BYTES[id(desc)] = ws_bytes
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        ins = SplatInputs(_ptr(pts), _ptr(points_int), _ptr(means), _ptr(means_int), _ptr(opa), _ptr(sem), _ptr(cov),
                          _ptr(radii), _ptr(scales))
        outs = SplatOutputs(_ptr(logits), _ptr(binl), _ptr(dens), _ptr(probability), _ptr(argmax_out))
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(L.gf_splat_forward(ctypes.byref(desc), ctypes.byref(ins), ctypes.byref(outs), _ptr(ws), ws_bytes,
                                      stream))
    return (logits, binl, dens, probability), ws


def splat_backward_raw(desc, pts, means, opa, sem, cov, grads_in, saved, *, points_int=None, means_int=None,
                       radii=None, scales=None):
    """One sample through ``gf_splat_backward``.  Returns (g_means[G,3], g_opa[G], g_sem[G,C], g_cov[G,cov_stride])."""
    L = _lib.lib()
    dev = pts.device
    with torch.cuda.device(dev):
        G, C = desc.G, desc.C
        gm = torch.empty((G, 3), dtype=torch.float32, device=dev)
        go = torch.empty((G,), dtype=torch.float32, device=dev)
        gs = torch.empty((G, C), dtype=torch.float32, device=dev)
        gc = torch.empty((G, desc.cov_stride), dtype=torch.float32, device=dev)
        ws_bytes = L.gf_splat_backward_workspace_bytes(ctypes.byref(desc))
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
        ins = SplatInputs(_ptr(pts), _ptr(points_int), _ptr(means), _ptr(means_int), _ptr(opa), _ptr(sem), _ptr(cov),
                          _ptr(radii), _ptr(scales))
        g_logits, g_bin, g_dens = grads_in
        logits, binl, probability = saved
        gr = SplatGrads(_ptr(g_logits), _ptr(g_bin), _ptr(g_dens), _ptr(logits), _ptr(binl), _ptr(probability),
                        _ptr(gm), _ptr(go), _ptr(gs), _ptr(gc))
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(L.gf_splat_backward(ctypes.byref(desc), ctypes.byref(ins), ctypes.byref(gr), _ptr(ws), ws_bytes,
                                       stream))
    return gm, go, gs, gc


def read_flags(ws, device):
    flags = ctypes.c_uint32(0)
    with torch.cuda.device(device):
        stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        _lib.check(_lib.lib().gf_splat_read_flags(_ptr(ws), stream, ctypes.byref(flags)))
    return int(flags.value)


def _assert_flags(flags):
    # same conditions, same exception type as the reference's Python asserts
    assert not (flags & _lib.GF_FLAG_POINT_OUT_OF_GRID), "points_int outside the voxel grid"
    assert not (flags & _lib.GF_FLAG_MEAN_OUT_OF_GRID), "means3D_int outside the voxel grid"
    assert not (flags & _lib.GF_FLAG_RADIUS_LT_1), "radii.min() < 1"


class _SplatFunction(torch.autograd.Function):
    """Per-sample autograd bridge (reference: ``_LocalAggregate``, ``__init__.py:18-106``).

    Saves only the user tensors (+ the prob outputs); the backward kernels need no scratch kept
    alive from the forward, unlike the reference's three byte buffers (``__init__.py:53-63``).
    """

    @staticmethod
    def forward(ctx, pts, means, opa, sem, scales, cov, cfg):
        args = (means.shape[0], pts.shape[0], sem.shape[1], cfg["H"], cfg["W"], cfg["D"], cfg["variant"],
                cfg["radii_axes"], 9, cfg["pc_min"], cfg["grid_size"], cfg["scale_multiplier"], cfg["radii_min"])
        desc = _cached_desc(args, *args)
        pts_c, means_c, opa_c, sem_c, scales_c = map(_f32c, (pts, means, opa, sem, scales))
        cov_c = _f32c(cov).reshape(-1, 9)
        (logits, binl, dens, probability), ws = splat_forward_raw(desc, pts_c, means_c, opa_c, sem_c, cov_c,
                                                                 scales=scales_c)
        if cfg["validate"]:
            _assert_flags(read_flags(ws, pts.device))
        ctx.desc = desc
        ctx.prob = cfg["variant"] == _lib.GF_SPLAT_PROB
        if ctx.prob:
            ctx.save_for_backward(pts_c, means_c, opa_c, sem_c, scales_c, cov_c, logits, binl, probability)
            ctx.mark_non_differentiable(probability)
            return logits, binl, dens, probability
        ctx.save_for_backward(pts_c, means_c, opa_c, sem_c, scales_c, cov_c)
        return logits

    @staticmethod
    def backward(ctx, *grad_outputs):
        if ctx.prob:
            pts, means, opa, sem, scales, cov, logits, binl, probability = ctx.saved_tensors
            g_logits, g_bin, g_dens = (None if g is None else _f32c(g) for g in grad_outputs[:3])
            N, C = logits.shape
            if g_logits is None:
                g_logits = torch.zeros((N, C), dtype=torch.float32, device=pts.device)
            if g_bin is None:
                g_bin = torch.zeros((N,), dtype=torch.float32, device=pts.device)
            if g_dens is None:
                g_dens = torch.zeros((N,), dtype=torch.float32, device=pts.device)
            grads_in, saved = (g_logits, g_bin, g_dens), (logits, binl, probability)
        else:
            pts, means, opa, sem, scales, cov = ctx.saved_tensors
            grads_in, saved = (_f32c(grad_outputs[0]), None, None), (None, None, None)
        # cov_stride is 9 here: the kernel writes the gradient straight into the 3x3 layout (the six gathered
        # entries carry it, the lower triangle is zero -- indexing autograd in the reference:
        # cov3D.flatten(1)[:, [0,4,8,1,5,2]])
        gm, go, gs, gcov = splat_backward_raw(ctx.desc, pts, means, opa, sem, cov, grads_in, saved, scales=scales)
        return None, gm, go, gs, None, gcov.view(-1, 3, 3), None


class _LocalAggregatorBase(nn.Module):
    _variant = _lib.GF_SPLAT_BASE
    _radii_axes = 1

    def _setup(self, scale_multiplier, H, W, D, pc_min, grid_size, radii_min):
        self.scale_multiplier = scale_multiplier
        self.H, self.W, self.D = H, W, D
        self.register_buffer("pc_min", torch.tensor(pc_min, dtype=torch.float).unsqueeze(0))
        self._pc_min_host = tuple(float(v) for v in pc_min)
        self.grid_size = grid_size
        self.radii_min = radii_min
        #: read the device status word after each forward and raise AssertionError like the
        #: reference's asserts (one sync).  Set False for fully asynchronous / graph-captured use.
        self.validate = os.environ.get("GF_B200_VALIDATE", "1") != "0"
        _lib.lib()  # fail loudly at construction time if the extension is missing

    def _cfg(self):
        cfg = self.__dict__.get("_cfg_cache")
        if cfg is None or cfg["validate"] != self.validate:
            cfg = dict(H=self.H, W=self.W, D=self.D, variant=self._variant, radii_axes=self._radii_axes,
                       pc_min=self._pc_min_host, grid_size=float(self.grid_size),
                       scale_multiplier=float(self.scale_multiplier),
                       radii_min=int(self.radii_min) if self.radii_min is not None else 0, validate=self.validate)
            self.__dict__["_cfg_cache"] = cfg
        return cfg

    def forward_from_srt(self, pts, means3D, opacities, semantics, scales, rotations):
        """Entry point next to the reference signature: takes the Gaussians' scales and
        rotation quaternions instead of a precomputed inverse covariance and never leaves the device."""
        return self.forward(pts, means3D, opacities, semantics, scales, inverse_covariance_from_srt(scales, rotations))

    def grid_points(self, device):
        """``[1, H*W*D, 3]`` voxel-centre coordinates of this aggregator's grid on ``device``, built once with the
        reference loader's arithmetic (``LoadOccupancySurroundOcc.get_meshgrid``, dataset/transform_3d.py:487-499:
        ``arange * reso + 0.5 * reso + min`` in fp32, x-major) — the ``occ_xyz`` every shipped config feeds as ``pts``."""
        cache = self.__dict__.setdefault("_grid_pts", {})
        key = str(device)
        if key not in cache:
            from .synthetic import voxel_centers
            cache[key] = voxel_centers((self.H, self.W, self.D), self._pc_min_host, float(self.grid_size)).reshape(1, -1, 3).to(device)
        return cache[key]

    def forward_on_grid(self, means3D, opacities, semantics, scales, cov3D):
        """Entry point next to the reference signature: ``forward(pts = the grid's own voxel centres, ...)`` with the
        points kept resident on the device (a caller that evaluates on the occupancy grid itself, as all shipped
        configs do, then ships only the Gaussians).  Same kernels, same results as passing ``occ_xyz``."""
        return self.forward(self.grid_points(means3D.device), means3D, opacities, semantics, scales, cov3D)

    def _run(self, pts, means3D, opacities, semantics, scales, cov3D):
        _require_cuda(pts, means3D, opacities, semantics, scales, cov3D)
        assert not pts.requires_grad
        B = pts.shape[0]
        cfg = self._cfg()
        outs = []
        for b in range(B):
            outs.append(_SplatFunction.apply(pts[b], means3D[b], opacities[b], semantics[b],
                                             scales[b].detach(), cov3D[b], cfg))
        return outs


def inverse_covariance_from_srt(scales, rotations):
    """Sigma^-1 = R^T diag(1/s^2) R on the device, differentiable.

    ``GaussianHead.prepare_gaussian_args`` (model/head/gaussian_head.py:111-119) builds ``Cov = (S R)^T (S R)``
    and inverts it numerically on the CPU (``Cov.cpu().inverse().cuda()``, a blocking round trip per
    supervised layer).  ``R`` is a rotation, so the inverse is available in closed form; the quaternion ->
    matrix map is the reference's (model/utils/utils.py:20-66, (w, x, y, z), normalised first)."""
    q = torch.nn.functional.normalize(rotations, dim=-1)
    w, x, y, z = q.unbind(-1)
    R = torch.stack([
        torch.stack([w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
        torch.stack([2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)], -1),
        torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z], -1),
    ], dim=-2)
    inv_s2 = 1.0 / (scales * scales)
    return torch.matmul(R.transpose(-1, -2) * inv_s2.unsqueeze(-2), R)


class LocalAggregator(_LocalAggregatorBase):
    """Drop-in for ``local_aggregate.LocalAggregator``: forward(...) -> logits [N, C]."""

    def __init__(self, scale_multiplier, H, W, D, pc_min, grid_size, inv_softmax=False):
        super().__init__()
        self._setup(scale_multiplier, H, W, D, pc_min, grid_size, radii_min=None)
        self.inv_softmax = inv_softmax

    def forward(self, pts, means3D, opacities, semantics, scales, cov3D):
        outs = self._run(pts, means3D, opacities, semantics, scales, cov3D)
        assert not self.inv_softmax  # the reference's `assert False` branch
        return outs[0] if len(outs) == 1 else torch.stack(outs, 0)


    @torch.no_grad()
    def forward_with_occupancy(self, pts, means3D, opacities, semantics, scales, cov3D):
        """Inference helper beyond the reference API: returns ``(logits [N,C], occ [N] uint8)`` with
        ``occ == logits.argmax(1)`` computed in the render epilogue (what ``GaussianHead.forward`` does
        next, model/head/gaussian_head.py:185), so the logits are not re-read.  Batch of 1."""
        _require_cuda(pts, means3D, opacities, semantics, scales, cov3D)
        assert pts.shape[0] == 1
        cfg = self._cfg()
        args = (means3D.shape[1], pts.shape[1], semantics.shape[2], cfg["H"], cfg["W"], cfg["D"], cfg["variant"],
                cfg["radii_axes"], 9, cfg["pc_min"], cfg["grid_size"], cfg["scale_multiplier"], cfg["radii_min"])
        desc = _cached_desc(args, *args)
        occ = torch.empty(pts.shape[1], dtype=torch.uint8, device=pts.device)
        (logits, _, _, _), ws = splat_forward_raw(desc, _f32c(pts[0]), _f32c(means3D[0]), _f32c(opacities[0]),
                                                 _f32c(semantics[0]), _f32c(cov3D[0]).reshape(-1, 9),
                                                 scales=_f32c(scales[0]), argmax_out=occ)
        if self.validate:
            _assert_flags(read_flags(ws, pts.device))
        return logits, occ


class LocalAggregatorProb(_LocalAggregatorBase):
    """Drop-in for ``local_aggregate_prob.LocalAggregator``: -> (logits [N,C], bin_logits [N], density [N])."""
    _variant = _lib.GF_SPLAT_PROB

    def __init__(self, scale_multiplier, H, W, D, pc_min, grid_size, radii_min=1):
        super().__init__()
        self._setup(scale_multiplier, H, W, D, pc_min, grid_size, radii_min=radii_min)

    def forward(self, pts, means3D, opas, semantics, scales, cov3D):
        outs = self._run(pts, means3D, opas, semantics, scales, cov3D)
        if len(outs) == 1:
            return outs[0][0], outs[0][1], outs[0][2]
        return tuple(torch.stack([o[i] for o in outs], 0) for i in range(3))


class LocalAggregatorProbFast(LocalAggregatorProb):
    """Drop-in for ``local_aggregate_prob_fast.LocalAggregator`` (per-axis integer radii)."""
    _radii_axes = 3
