"""Drop-in for ``model/encoder/gaussian_encoder/ops/deformable_aggregation.py:7-117``.

``DeformableAggregationFunction.apply(mc_ms_feat, spatial_shape, scale_start_index,
sampling_location, weights)`` and the static ``feature_maps_format`` keep the reference's
signatures and dtype normalisation (``.float()`` / ``.int()``); the kernels are the sm_100a ones
behind ``gf_daf_forward`` / ``gf_daf_backward`` (``include/gf_b200.h``).

``deformable_aggregation_fused`` is an opt-in entry point next to it: the op together
with the masked joint softmax of its weights and the sum over the key points that
``DeformableFeatureAggregation.forward`` wraps around it (``deformable_module.py:213-228,242``).
"""
from __future__ import annotations

import ctypes

import torch
from torch.autograd.function import Function, once_differentiable

from .. import _lib
from .._lib import DAF_MAX_LEVELS, DafDesc, DafFormatDesc, DafFusedDesc


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _desc(feat, spatial_shape, sampling_location, weights):
    d = DafDesc()
    d.batch, d.num_cams, d.num_feat, d.num_embeds = feat.shape
    d.num_scale = spatial_shape.shape[0]
    d.num_pts = sampling_location.shape[1]
    d.num_groups = weights.shape[4]
    return d


def _a16(t):
    """The kernels read / write / atomically add 16-byte vectors: a tensor whose storage offset leaves it misaligned
    (a slice of a larger buffer) is copied to a fresh allocation; the C ABI rejects misaligned pointers."""
    return t if t.data_ptr() % 16 == 0 else t.clone(memory_format=torch.contiguous_format)


def _normalise(mc_ms_feat, spatial_shape, scale_start_index, sampling_location, weights):
    for t in (mc_ms_feat, spatial_shape, scale_start_index, sampling_location, weights):
        if not t.is_cuda:
            raise RuntimeError("DeformableAggregationFunction is CUDA-only (sm_100a); there is no CPU fallback.")
    return (_a16(mc_ms_feat.contiguous().float()), spatial_shape.contiguous().int(),
            scale_start_index.contiguous().int(), _a16(sampling_location.contiguous().float()),
            _a16(weights.contiguous().float()))


def _format_call(maps, table, channels, inverse):
    d = DafFormatDesc()
    d.batch_cams = maps[0].shape[0] * maps[0].shape[1]
    d.num_embeds = channels
    d.num_scale = len(maps)
    for i, fm in enumerate(maps):
        d.hw[i] = fm.shape[-2] * fm.shape[-1]
    ptrs = (ctypes.c_void_p * len(maps))(*[fm.data_ptr() for fm in maps])
    dev = table.device
    with torch.cuda.device(dev):
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(_lib.lib().gf_daf_format(ctypes.byref(d), ptrs, _ptr(table), int(inverse), stream))


_LEVEL_TENSORS = {}


def _level_tensors(shapes, starts, dev):
    """The two small int64 tensors the reference rebuilds (and copies host->device) on every call; they depend
    on the level geometry only, so one copy per (geometry, device) is kept."""
    key = (tuple(shapes), str(dev))
    hit = _LEVEL_TENSORS.get(key)
    if hit is None:
        hit = (torch.tensor(shapes, dtype=torch.int64, device=dev), torch.tensor(starts, dtype=torch.int64, device=dev))
        if len(_LEVEL_TENSORS) > 64:
            _LEVEL_TENSORS.clear()
        _LEVEL_TENSORS[key] = hit
    return hit


class _FeatureMapsToTable(Function):
    """``[B, M, C, h_l, w_l]`` maps -> contiguous channels-last table ``[B, M, sum(h_l*w_l), C]`` in one pass
    (``gf_daf_format``); the gradient is the same kernel run in the other direction."""

    @staticmethod
    def forward(ctx, *maps):
        maps = [_a16(fm.contiguous()) for fm in maps]
        bs, num_cams, channels = maps[0].shape[:3]
        ctx.shapes = [tuple(fm.shape) for fm in maps]
        total = sum(fm.shape[-2] * fm.shape[-1] for fm in maps)
        table = torch.empty((bs, num_cams, total, channels), dtype=torch.float32, device=maps[0].device)
        _format_call(maps, table, channels, False)
        return table

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_table):
        grad_table = _a16(grad_table.contiguous().float())
        grads = [torch.empty(shape, dtype=torch.float32, device=grad_table.device) for shape in ctx.shapes]
        _format_call(grads, grad_table, ctx.shapes[0][2], True)
        return tuple(grads)


def feature_maps_format(feature_maps, inverse=False):
    """List of ``[B, M, C, h_l, w_l]`` maps <-> (``[B, M, sum(h_l*w_l), C]``, shapes, start offsets)
    — ``ops/deformable_aggregation.py:78-117``.  For float32 CUDA maps the table is produced contiguous by one
    fused transpose kernel (the reference returns a permuted view and pays a transposing copy in every later
    ``.contiguous()``); other inputs take the reference's reshape/cat/permute route."""
    if not inverse:
        bs, num_cams, channels = feature_maps[0].shape[:3]
        shapes, starts, total = [], [], 0
        for fm in feature_maps:
            h, w = fm.shape[-2:]
            shapes.append((h, w))
            starts.append(total)
            total += h * w
        fused = (len(feature_maps) <= DAF_MAX_LEVELS and
                 all(fm.is_cuda and fm.dtype == torch.float32 and fm.dim() == 5 and fm.shape[:3] == (bs, num_cams, channels)
                     for fm in feature_maps))
        if fused:
            col = _FeatureMapsToTable.apply(*feature_maps)
        else:
            flat = [fm.reshape(bs, num_cams, channels, -1) for fm in feature_maps]
            col = torch.cat(flat, dim=-1).permute(0, 1, 3, 2)
        shape_t, start_t = _level_tensors(shapes, starts, col.device)
        return [col, shape_t, start_t]
    col, spatial_shape = feature_maps[0], feature_maps[1].int()
    sizes = (spatial_shape[:, 0] * spatial_shape[:, 1]).tolist()
    chunks = torch.split(col.permute(0, 1, 3, 2), sizes, dim=-1)
    return [c.reshape(c.shape[:3] + (int(spatial_shape[i, 0]), int(spatial_shape[i, 1])))
            for i, c in enumerate(chunks)]


class DeformableAggregationFunction(Function):
    @staticmethod
    def forward(ctx, mc_ms_feat, spatial_shape, scale_start_index, sampling_location, weights):
        mc_ms_feat, spatial_shape, scale_start_index, sampling_location, weights = _normalise(
            mc_ms_feat, spatial_shape, scale_start_index, sampling_location, weights)
        d = _desc(mc_ms_feat, spatial_shape, sampling_location, weights)
        dev = mc_ms_feat.device
        with torch.cuda.device(dev):
            output = torch.empty((d.batch, d.num_pts, d.num_embeds), dtype=torch.float32, device=dev)
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.lib().gf_daf_forward(ctypes.byref(d), _ptr(mc_ms_feat), _ptr(spatial_shape),
                                                 _ptr(scale_start_index), _ptr(sampling_location), _ptr(weights),
                                                 _ptr(output), stream))
        ctx.save_for_backward(mc_ms_feat, spatial_shape, scale_start_index, sampling_location, weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        mc_ms_feat, spatial_shape, scale_start_index, sampling_location, weights = ctx.saved_tensors
        d = _desc(mc_ms_feat, spatial_shape, sampling_location, weights)
        dev = mc_ms_feat.device
        grad_output = _a16(grad_output.contiguous().float())
        with torch.cuda.device(dev):
            grad_feat = torch.zeros_like(mc_ms_feat)
            grad_loc = torch.zeros_like(sampling_location)
            grad_w = torch.zeros_like(weights)
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.lib().gf_daf_backward(ctypes.byref(d), _ptr(mc_ms_feat), _ptr(spatial_shape),
                                                  _ptr(scale_start_index), _ptr(sampling_location), _ptr(weights),
                                                  _ptr(grad_output), _ptr(grad_feat), _ptr(grad_loc), _ptr(grad_w),
                                                  stream))
        return grad_feat, None, None, grad_loc, grad_w

    feature_maps_format = staticmethod(feature_maps_format)


class DeformableAggregationFusedFunction(Function):
    """``[B, A, C] = sum_k DAF(feat, loc, masked_softmax(weight_logits))[b, a*K + k]`` in one kernel
    (``gf_daf_fused_forward`` / ``gf_daf_fused_backward``).  Gradients: ``mc_ms_feat``,
    ``sampling_location``, ``weight_logits``."""

    @staticmethod
    def forward(ctx, mc_ms_feat, spatial_shape, scale_start_index, sampling_location, weight_logits,
                point_mask=None, weight_mask=None):
        if weight_logits.dim() != 6:
            raise ValueError("weight_logits must be [B, A, K, M, L, Gr] (deformable_module.py:177-189)")
        B, A, K, M, L, Gr = weight_logits.shape
        mc_ms_feat, spatial_shape, scale_start_index, sampling_location, logits = _normalise(
            mc_ms_feat, spatial_shape, scale_start_index, sampling_location, weight_logits)
        if sampling_location.shape != (B, A * K, M, 2):
            raise ValueError(f"sampling_location must be [B, A*K, M, 2] = {(B, A * K, M, 2)}, got {tuple(sampling_location.shape)}")
        if mc_ms_feat.dim() != 4 or mc_ms_feat.shape[0] != B or mc_ms_feat.shape[1] != M or spatial_shape.shape[0] != L:
            raise ValueError(f"feature table [B, M, F, C] / level table do not match the weights: feat {tuple(mc_ms_feat.shape)}, "
                             f"levels {spatial_shape.shape[0]}, weights B={B} M={M} L={L}")
        masks = []
        for name, mk, shape in (("point_mask", point_mask, (B, A, K, M)), ("weight_mask", weight_mask, (B, A, K, M, L, Gr))):
            if mk is not None:
                if not mk.is_cuda or tuple(mk.shape) != shape:
                    raise ValueError(f"{name} must be a CUDA tensor of shape {shape}")
                mk = (mk if mk.dtype in (torch.bool, torch.uint8) else mk != 0).contiguous()
            masks.append(mk)
        point_mask, weight_mask = masks
        fd = DafFusedDesc()
        fd.d = _desc(mc_ms_feat, spatial_shape, sampling_location, logits.view(B, A * K, M, L, Gr))
        fd.pts_per_anchor = K
        dev = mc_ms_feat.device
        with torch.cuda.device(dev):
            output = torch.empty((B, A, fd.d.num_embeds), dtype=torch.float32, device=dev)
            stats = torch.empty((B, A, Gr, 2), dtype=torch.float32, device=dev)
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.lib().gf_daf_fused_forward(
                ctypes.byref(fd), _ptr(mc_ms_feat), _ptr(spatial_shape), _ptr(scale_start_index), _ptr(sampling_location),
                _ptr(logits), _ptr(point_mask) if point_mask is not None else None,
                _ptr(weight_mask) if weight_mask is not None else None, _ptr(output), _ptr(stats), stream))
        ctx.fd = fd
        ctx.has_masks = (point_mask is not None, weight_mask is not None)
        saved = [mc_ms_feat, spatial_shape, scale_start_index, sampling_location, logits, stats, output]
        saved += [m for m in (point_mask, weight_mask) if m is not None]
        ctx.save_for_backward(*saved)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        saved = list(ctx.saved_tensors)
        mc_ms_feat, spatial_shape, scale_start_index, sampling_location, logits, stats, output = saved[:7]
        rest = saved[7:]
        point_mask = rest.pop(0) if ctx.has_masks[0] else None
        weight_mask = rest.pop(0) if ctx.has_masks[1] else None
        dev = mc_ms_feat.device
        grad_output = _a16(grad_output.contiguous().float())
        with torch.cuda.device(dev):
            grad_feat = torch.zeros_like(mc_ms_feat)
            grad_loc = torch.zeros_like(sampling_location)
            grad_logits = torch.empty_like(logits)
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.lib().gf_daf_fused_backward(
                ctypes.byref(ctx.fd), _ptr(mc_ms_feat), _ptr(spatial_shape), _ptr(scale_start_index),
                _ptr(sampling_location), _ptr(logits), _ptr(point_mask) if point_mask is not None else None,
                _ptr(weight_mask) if weight_mask is not None else None, _ptr(stats), _ptr(output), _ptr(grad_output),
                _ptr(grad_feat), _ptr(grad_loc), _ptr(grad_logits), stream))
        return grad_feat, None, None, grad_loc, grad_logits, None, None


def fused_supported(num_embeds, num_groups, num_cams, num_levels, num_feat=1, num_pts=1):
    """True when ``deformable_aggregation_fused`` has a kernel for this shape (otherwise compose
    ``DeformableAggregationFunction`` with the PyTorch softmax as the reference does)."""
    fd = DafFusedDesc()
    fd.d.batch, fd.d.num_cams, fd.d.num_feat, fd.d.num_embeds = 1, num_cams, num_feat, num_embeds
    fd.d.num_scale, fd.d.num_pts, fd.d.num_groups = num_levels, num_pts, num_groups
    fd.pts_per_anchor = num_pts
    return bool(_lib.lib().gf_daf_fused_supported(ctypes.byref(fd)))


def deformable_aggregation_fused(mc_ms_feat, spatial_shape, scale_start_index, sampling_location, weight_logits,
                                 point_mask=None, weight_mask=None):
    """Replacement for ``deformable_module.py:213-228`` + ``:242`` around the op call:

    * ``weight_logits`` ``[B, A, K, M, L, Gr]`` — the raw weights after the permute at ``:177-189``
    * ``point_mask`` ``[B, A, K, M]`` bool — ``mask.permute(0, 2, 3, 1)`` of ``project_points`` (``:211``)
    * ``weight_mask`` ``[B, A, K, M, L, Gr]`` bool — the attn-drop mask (``:190-202``), or None in eval
    * ``sampling_location`` ``[B, A*K, M, 2]`` — ``points_2d`` as passed to the op

    Returns ``features.sum(dim=2)``: ``[B, A, C]``."""
    return DeformableAggregationFusedFunction.apply(mc_ms_feat, spatial_shape, scale_start_index, sampling_location,
                                                    weight_logits, point_mask, weight_mask)
