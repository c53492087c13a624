"""Drop-in for ``model/encoder/gaussian_encoder/ops/deformable_aggregation.py:7-117``.

``DeformableAggregationFunction.apply(mc_ms_feat, spatial_shape, scale_start_index,
sampling_location, weights)`` and the static ``feature_maps_format`` keep the reference's
signatures and dtype normalisation (``.float()`` / ``.int()``); the kernels are the sm_100a ones
behind ``gf_daf_forward`` / ``gf_daf_backward`` (``include/gf_b200.h``).
"""
from __future__ import annotations

import ctypes

import torch
from torch.autograd.function import Function, once_differentiable

from .. import _lib
from .._lib import DAF_MAX_LEVELS, DafDesc, DafFormatDesc


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _desc(feat, spatial_shape, sampling_location, weights):
    d = DafDesc()
    d.batch, d.num_cams, d.num_feat, d.num_embeds = feat.shape
    d.num_scale = spatial_shape.shape[0]
    d.num_pts = sampling_location.shape[1]
    d.num_groups = weights.shape[4]
    return d


def _normalise(mc_ms_feat, spatial_shape, scale_start_index, sampling_location, weights):
    for t in (mc_ms_feat, spatial_shape, scale_start_index, sampling_location, weights):
        if not t.is_cuda:
            raise RuntimeError("DeformableAggregationFunction is CUDA-only (sm_100a); there is no CPU fallback.")
    return (mc_ms_feat.contiguous().float(), spatial_shape.contiguous().int(),
            scale_start_index.contiguous().int(), sampling_location.contiguous().float(),
            weights.contiguous().float())


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
        maps = [fm.contiguous() for fm in maps]
        bs, num_cams, channels = maps[0].shape[:3]
        ctx.shapes = [tuple(fm.shape) for fm in maps]
        total = sum(fm.shape[-2] * fm.shape[-1] for fm in maps)
        table = torch.empty((bs, num_cams, total, channels), dtype=torch.float32, device=maps[0].device)
        _format_call(maps, table, channels, False)
        return table

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_table):
        grad_table = grad_table.contiguous().float()
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
        grad_output = grad_output.contiguous().float()
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
