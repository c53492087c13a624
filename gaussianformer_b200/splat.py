"""Host-side mirror of the reference's three splat packages.

``LocalAggregator`` keeps the reference's constructor kwargs, buffer name (``pc_min``) and call
signature (``model/head/localagg/local_aggregate/__init__.py:108-161``,
``model/head/localagg_prob/local_aggregate_prob/__init__.py:118-169``,
``model/head/localagg_prob_fast/local_aggregate_prob_fast/__init__.py:151``) so that
``GaussianHead`` (``model/head/gaussian_head.py:30-39,157-163``) can construct and call it
unchanged.  The arithmetic — including the reference's Python-side preparation (voxel indices,
radii, 3x3 -> 6 gather) — runs in the sm_100a kernels behind the C ABI (``include/gf_b200.h``).

Differences from the reference, all supersets:

* batch sizes > 1 are accepted (the reference asserts ``B == 1``) and run as ONE batched launch per kernel
  (``gf_splat_desc.batch``); ``B == 1`` returns the same squeezed shapes as the reference;
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
_WS_BYTES = {}   # id(cached desc) -> (forward workspace bytes, backward workspace bytes or None)


def _cached_desc(key, *args):
    """ctypes descriptors are immutable after construction here; build each distinct one once."""
    d = _DESC_CACHE.get(key)
    if d is None:
        d = _make_desc(*args)
        if len(_DESC_CACHE) < 256:
            _DESC_CACHE[key] = d
    return d


def _make_desc(G, N, C, H, W, D, variant, radii_axes, cov_stride, pc_min, grid_size, scale_multiplier, radii_min,
               batch=1, pts_shared=0):
    d = SplatDesc()
    d.G, d.N, d.C, d.H, d.W, d.D = G, N, C, H, W, D
    d.variant, d.radii_axes, d.cov_stride = variant, radii_axes, cov_stride
    d.pc_min[0], d.pc_min[1], d.pc_min[2] = pc_min
    d.grid_size, d.scale_multiplier, d.radii_min = grid_size, scale_multiplier, radii_min
    d.batch, d.pts_shared = batch, pts_shared
    return d


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "gaussianformer_b200 ops are CUDA-only (sm_100a); got a CPU tensor. There is no CPU fallback.")


def _forward_ws_bytes(desc):
    L = _lib.lib()
    hit = _WS_BYTES.get(id(desc))
    if hit is not None:
        return hit
    ws_bytes = L.gf_splat_forward_workspace_bytes(ctypes.byref(desc))
    if ws_bytes == 0:
        raise _lib.GfError(L.gf_last_error().decode())
    if any(desc is v for v in _DESC_CACHE.values()):
        _WS_BYTES[id(desc)] = ws_bytes
    return ws_bytes


def splat_forward_raw(desc, pts, means, opa, sem, cov, *, points_int=None, means_int=None, radii=None, scales=None,
                      rotations=None, argmax_out=None, logits_cn_out=None, labels=None, class_weights=None,
                      ce_partials_out=None, want_logits=True):
    """``desc.batch`` samples through ONE ``gf_splat_forward`` call.  Tensors carry the leading batch dimension
    (``pts`` may be ``[N,3]`` with ``desc.pts_shared``).  Returns ((logits, bin, density, probability), workspace)."""
    L = _lib.lib()
    dev = means.device
    B = max(int(desc.batch), 1)
    with torch.cuda.device(dev):
        N, C = desc.N, desc.C
        logits = torch.empty((B, N, C), dtype=torch.float32, device=dev) if want_logits else None
        prob = desc.variant == _lib.GF_SPLAT_PROB
        if prob:   # three independent tensors, like the reference returns (an in-place edit of one must not touch the others)
            binl = torch.empty((B, N), dtype=torch.float32, device=dev)
            dens = torch.empty((B, N), dtype=torch.float32, device=dev)
            probability = torch.empty((B, N), dtype=torch.float32, device=dev)
        else:
            binl = dens = probability = None
        ws_bytes = _forward_ws_bytes(desc)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        ins = SplatInputs(_ptr(pts), _ptr(points_int), _ptr(means), _ptr(means_int), _ptr(opa), _ptr(sem), _ptr(cov),
                          _ptr(radii), _ptr(scales), _ptr(rotations))
        outs = SplatOutputs(_ptr(logits), _ptr(binl), _ptr(dens), _ptr(probability), _ptr(argmax_out),
                            _ptr(logits_cn_out), _ptr(labels), _ptr(class_weights), _ptr(ce_partials_out))
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(L.gf_splat_forward(ctypes.byref(desc), ctypes.byref(ins), ctypes.byref(outs), _ptr(ws), ws_bytes,
                                      stream))
    return (logits, binl, dens, probability), ws


def splat_backward_raw(desc, pts, means, opa, sem, cov, grads_in, saved, *, points_int=None, means_int=None,
                       radii=None, scales=None, rotations=None):
    """``desc.batch`` samples through ONE ``gf_splat_backward`` call.  Returns (g_means [B,G,3], g_opa [B,G],
    g_sem [B,G,C], g_cov [B,G,cov_stride] or None, g_scales [B,G,3] or None, g_rotations [B,G,4] or None)."""
    L = _lib.lib()
    dev = means.device
    B = max(int(desc.batch), 1)
    with torch.cuda.device(dev):
        G, C = desc.G, desc.C
        gm = torch.empty((B, G, 3), dtype=torch.float32, device=dev)
        go = torch.empty((B, G), dtype=torch.float32, device=dev)
        gs = torch.empty((B, G, C), dtype=torch.float32, device=dev)
        srt = cov is None
        gc = None if srt else torch.empty((B, G, desc.cov_stride), dtype=torch.float32, device=dev)
        gsc = torch.empty((B, G, 3), dtype=torch.float32, device=dev) if srt else None
        grot = torch.empty((B, G, 4), dtype=torch.float32, device=dev) if srt else None
        ws_bytes = L.gf_splat_backward_workspace_bytes(ctypes.byref(desc))
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
        ins = SplatInputs(_ptr(pts), _ptr(points_int), _ptr(means), _ptr(means_int), _ptr(opa), _ptr(sem), _ptr(cov),
                          _ptr(radii), _ptr(scales), _ptr(rotations))
        g_logits, g_bin, g_dens = grads_in
        logits, binl, probability = saved
        gr = SplatGrads(_ptr(g_logits), _ptr(g_bin), _ptr(g_dens), _ptr(logits), _ptr(binl), _ptr(probability),
                        _ptr(gm), _ptr(go), _ptr(gs), _ptr(gc), _ptr(gsc), _ptr(grot))
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(L.gf_splat_backward(ctypes.byref(desc), ctypes.byref(ins), ctypes.byref(gr), _ptr(ws), ws_bytes,
                                       stream))
    return gm, go, gs, gc, gsc, grot


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


def _desc_for(cfg, B, G, N, C, pts_shared, cov_stride=9):
    args = (G, N, C, cfg["H"], cfg["W"], cfg["D"], cfg["variant"], cfg["radii_axes"], cov_stride, cfg["pc_min"],
            cfg["grid_size"], cfg["scale_multiplier"], cfg["radii_min"], B, int(pts_shared))
    return _cached_desc(args, *args)


class _SplatFunction(torch.autograd.Function):
    """Batched autograd bridge (reference: ``_LocalAggregate``, ``__init__.py:18-106``, per sample).

    Inputs carry the batch dimension: ``pts [B,N,3]`` (or ``[N,3]`` shared by the batch), ``means [B,G,3]``,
    ``opa [B,G]``, ``sem [B,G,C]``, ``scales [B,G,3]`` and either ``cov [B,G,3,3]`` (``rot`` None) or ``rot [B,G,4]``
    (``cov`` None: the inverse covariance is built inside the pack kernel).  One C-ABI call per direction for the whole
    batch.  Saves only the user tensors (+ the prob outputs); the backward kernels need no scratch kept alive from the
    forward, unlike the reference's three byte buffers (``__init__.py:53-63``)."""

    @staticmethod
    def forward(ctx, pts, means, opa, sem, scales, cov, rot, cfg):
        B, G = means.shape[0], means.shape[1]
        pts_shared = pts.dim() == 2
        N, C = pts.shape[-2], sem.shape[-1]
        desc = _desc_for(cfg, B, G, N, C, pts_shared)
        pts_c, means_c, opa_c, sem_c, scales_c = map(_f32c, (pts, means, opa, sem, scales))
        cov_c = None if cov is None else _f32c(cov).reshape(B, G, 9)
        rot_c = None if rot is None else _f32c(rot)
        (logits, binl, dens, probability), ws = splat_forward_raw(desc, pts_c, means_c, opa_c, sem_c, cov_c,
                                                                 scales=scales_c, rotations=rot_c)
        if cfg["validate"]:
            _assert_flags(read_flags(ws, means.device))
        ctx.desc = desc
        ctx.prob = cfg["variant"] == _lib.GF_SPLAT_PROB
        ctx.srt = cov is None
        extra = (rot_c,) if ctx.srt else (cov_c,)
        if ctx.prob:
            ctx.save_for_backward(pts_c, means_c, opa_c, sem_c, scales_c, *extra, logits, binl, probability)
            ctx.mark_non_differentiable(probability)
            return logits, binl, dens, probability
        ctx.save_for_backward(pts_c, means_c, opa_c, sem_c, scales_c, *extra)
        return logits

    @staticmethod
    def backward(ctx, *grad_outputs):
        if ctx.prob:
            pts, means, opa, sem, scales, cr, logits, binl, probability = ctx.saved_tensors
            g_logits, g_bin, g_dens = (None if g is None else _f32c(g) for g in grad_outputs[:3])
            if g_logits is None:
                g_logits = torch.zeros_like(logits)
            if g_bin is None:
                g_bin = torch.zeros_like(binl)
            if g_dens is None:
                g_dens = torch.zeros_like(binl)
            grads_in, saved = (g_logits, g_bin, g_dens), (logits, binl, probability)
        else:
            pts, means, opa, sem, scales, cr = ctx.saved_tensors
            grads_in, saved = (_f32c(grad_outputs[0]), None, None), (None, None, None)
        # cov_stride is 9 here: the kernel writes the gradient straight into the 3x3 layout (the six gathered
        # entries carry it, the lower triangle is zero -- indexing autograd in the reference:
        # cov3D.flatten(1)[:, [0,4,8,1,5,2]])
        cov, rot = (None, cr) if ctx.srt else (cr, None)
        gm, go, gs, gcov, gsc, grot = splat_backward_raw(ctx.desc, pts, means, opa, sem, cov, grads_in, saved,
                                                         scales=scales, rotations=rot)
        B, G = means.shape[0], means.shape[1]
        return (None, gm, go, gs, gsc, None if gcov is None else gcov.view(B, G, 3, 3), grot, None)


class _LocalAggregatorBase(nn.Module):
    _variant = _lib.GF_SPLAT_BASE
    _radii_axes = 1

    def _setup(self, scale_multiplier, H, W, D, pc_min, grid_size, radii_min):
        self.scale_multiplier = scale_multiplier
        self.H, self.W, self.D = H, W, D
        self.register_buffer("pc_min", torch.tensor(pc_min, dtype=torch.float).unsqueeze(0))
        self.grid_size = grid_size
        self.radii_min = radii_min
        #: read the device status word after each forward and raise AssertionError like the
        #: reference's asserts (one sync).  Set False for fully asynchronous / graph-captured use.
        self.validate = os.environ.get("GF_B200_VALIDATE", "1") != "0"
        _lib.lib()  # fail loudly at construction time if the extension is missing

    def _pc_min_host(self):
        """The registered ``pc_min`` buffer as host floats.  The reference reads the buffer on every call
        (``__init__.py:137``); here it is read back once per buffer version (``load_state_dict`` or an in-place edit
        bumps the version), so a checkpoint that carries a different origin is honoured without a sync per call."""
        buf = self.pc_min
        key = (id(buf), buf._version, buf.device)
        hit = self.__dict__.get("_pc_min_cache")
        if hit is None or hit[0] != key:
            hit = (key, tuple(float(v) for v in buf.detach().reshape(-1).cpu().tolist()))
            self.__dict__["_pc_min_cache"] = hit
        return hit[1]

    def _cfg(self):
        # rebuilt from the live attributes on every call (the reference reads them on every call too)
        return dict(H=int(self.H), W=int(self.W), D=int(self.D), variant=self._variant, radii_axes=self._radii_axes,
                    pc_min=self._pc_min_host(), grid_size=float(self.grid_size),
                    scale_multiplier=float(self.scale_multiplier),
                    radii_min=int(self.radii_min) if self.radii_min is not None else 0, validate=bool(self.validate))

    def forward_from_srt(self, pts, means3D, opacities, semantics, scales, rotations):
        """Entry point next to the reference signature: takes the Gaussians' scales and rotation
        quaternions instead of a precomputed inverse covariance.  Sigma^-1 = R^T diag(1/s^2) R is built inside the
        pack kernel (``gf_splat_inputs.rotations``) and the backward returns the gradients of ``scales`` (through
        Sigma^-1 only; the radii stay detached like in the reference) and ``rotations`` from its own small kernel --
        no PyTorch op, no ``Cov.cpu().inverse().cuda()`` round trip (model/head/gaussian_head.py:111-119)."""
        return self._finish(self._run(pts, means3D, opacities, semantics, scales, None, rotations))

    def grid_points(self, device):
        """``[1, H*W*D, 3]`` voxel-centre coordinates of this aggregator's grid on ``device``, built once with the
        reference loader's arithmetic (``LoadOccupancySurroundOcc.get_meshgrid``, dataset/transform_3d.py:487-499:
        ``arange * reso + 0.5 * reso + min`` in fp32, x-major) — the ``occ_xyz`` every shipped config feeds as ``pts``."""
        cache = self.__dict__.setdefault("_grid_pts", {})
        key = (str(device), self._pc_min_host(), float(self.grid_size), self.H, self.W, self.D)
        if key not in cache:
            from .synthetic import voxel_centers
            cache[key] = voxel_centers((self.H, self.W, self.D), self._pc_min_host(), float(self.grid_size)).reshape(1, -1, 3).to(device)
        return cache[key]

    def forward_on_grid(self, means3D, opacities, semantics, scales, cov3D):
        """Entry point next to the reference signature: ``forward(pts = the grid's own voxel centres, ...)`` with the
        points kept resident on the device and SHARED by the whole batch (``gf_splat_desc.pts_shared``) — a caller
        that evaluates on the occupancy grid itself, as all shipped configs do, then ships only the Gaussians.  Same
        kernels, same results as passing ``occ_xyz``."""
        return self._finish(self._run(self.grid_points(means3D.device)[0], means3D, opacities, semantics, scales, cov3D, None))

    def _run(self, pts, means3D, opacities, semantics, scales, cov3D, rotations):
        _require_cuda(pts, means3D, opacities, semantics, scales, cov3D, rotations)
        assert not pts.requires_grad
        cfg = self._cfg()
        sc = scales.detach() if rotations is None else scales    # reference: radii from detached scales (__init__.py:134)
        return _SplatFunction.apply(pts, means3D, opacities, semantics, sc, cov3D, rotations, cfg)

    def _finish(self, out):
        raise NotImplementedError


def inverse_covariance_from_srt(scales, rotations):
    """Sigma^-1 = R^T diag(1/s^2) R on the device, differentiable (PyTorch ops; the in-kernel route is
    ``forward_from_srt``).

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

    def _finish(self, logits):
        assert not self.inv_softmax  # the reference's `assert False` branch
        return logits[0] if logits.shape[0] == 1 else logits

    def forward(self, pts, means3D, opacities, semantics, scales, cov3D):
        return self._finish(self._run(pts, means3D, opacities, semantics, scales, cov3D, None))

    @torch.no_grad()
    def forward_with_occupancy(self, pts, means3D, opacities, semantics, scales, cov3D):
        """Inference helper beyond the reference API: returns ``(logits [N,C], occ [N] uint8)`` with
        ``occ == logits.argmax(1)`` computed in the render epilogue (what ``GaussianHead.forward`` does
        next, model/head/gaussian_head.py:185), so the logits are not re-read.  Batch of 1."""
        assert pts.shape[0] == 1
        r = self.forward_eval(pts, means3D, opacities, semantics, scales, cov3D, layout="nc")
        return r["logits"][0], r["final_occ"][0]

    @torch.no_grad()
    def forward_eval(self, pts, means3D, opacities, semantics, scales, cov3D, labels=None, class_weights=None,
                     layout="cn"):
        """Post-op fusion toward the loss, inference only.  One launch produces what
        ``GaussianHead.forward`` and the loss build from the logits afterwards:

        * ``pred_occ``  ``[B, C, N]`` -- the class-major layout ``semantics[None].transpose(1, 2)`` hands to the loss
          (gaussian_head.py:165-175), written directly by the render epilogue (``layout="cn"``); ``layout="nc"`` returns
          ``logits [B, N, C]`` instead and ``layout=None`` neither (the 46 MB of logits then never reach HBM);
        * ``final_occ`` ``[B, N]`` uint8 -- ``prediction.argmax(dim=1)`` (gaussian_head.py:185);
        * ``ce_loss``   scalar -- ``CE_ssc_loss(pred_occ, labels, class_weights, ignore_index=255)``
          (loss/occupancy_loss.py:113-127,164-178) from per-CTA partial sums of the epilogue, when ``labels``
          (``[B, N]`` uint8 / int64, 255 = ignore) is given.  Needs one point per voxel in grid order."""
        _require_cuda(pts, means3D, opacities, semantics, scales, cov3D)
        cfg = self._cfg()
        B, G = means3D.shape[0], means3D.shape[1]
        pts_shared = pts.dim() == 2
        N, C = pts.shape[-2], semantics.shape[-1]
        desc = _desc_for(cfg, B, G, N, C, pts_shared)
        dev = means3D.device
        occ = torch.empty((B, N), dtype=torch.uint8, device=dev)
        cn = torch.empty((B, C, N), dtype=torch.float32, device=dev) if layout == "cn" else None
        lab = cw = part = None
        if labels is not None:
            lab = labels.to(device=dev, dtype=torch.uint8).reshape(B, N).contiguous()
            cw = None if class_weights is None else _f32c(class_weights.to(dev))
            rows = _lib.lib().gf_splat_ce_partials(ctypes.byref(desc))
            if rows == 0:
                raise ValueError("the fused cross-entropy needs one point per voxel (N == H*W*D)")
            part = torch.empty((rows, 2), dtype=torch.float32, device=dev)
        (logits, _, _, _), ws = splat_forward_raw(desc, _f32c(pts), _f32c(means3D), _f32c(opacities), _f32c(semantics),
                                                 _f32c(cov3D).reshape(B, G, 9), scales=_f32c(scales), argmax_out=occ,
                                                 logits_cn_out=cn, labels=lab, class_weights=cw, ce_partials_out=part,
                                                 want_logits=layout == "nc")
        if self.validate:
            flags = read_flags(ws, dev)
            _assert_flags(flags)
            if part is not None and (flags & _lib.GF_FLAG_GENERIC_PATH):
                raise ValueError("the fused cross-entropy needs the points in the grid's own voxel order")
        out = {"final_occ": occ, "pred_occ": cn, "logits": logits, "ce_loss": None}
        if part is not None:
            s = part.double().sum(0)
            out["ce_loss"] = (s[0] / s[1]).float()
        return out


class LocalAggregatorProb(_LocalAggregatorBase):
    """Drop-in for ``local_aggregate_prob.LocalAggregator``: -> (logits [N,C], bin_logits [N], density [N])."""
    _variant = _lib.GF_SPLAT_PROB

    def __init__(self, scale_multiplier, H, W, D, pc_min, grid_size, radii_min=1):
        super().__init__()
        self._setup(scale_multiplier, H, W, D, pc_min, grid_size, radii_min=radii_min)

    def _finish(self, outs):
        lg, bl, de = outs[0], outs[1], outs[2]
        if lg.shape[0] == 1:
            return lg[0], bl[0], de[0]
        return lg, bl, de

    def forward(self, pts, means3D, opas, semantics, scales, cov3D):
        return self._finish(self._run(pts, means3D, opas, semantics, scales, cov3D, None))


class LocalAggregatorProbFast(LocalAggregatorProb):
    """Drop-in for ``local_aggregate_prob_fast.LocalAggregator`` (per-axis integer radii)."""
    _radii_axes = 3
