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
from .._lib import DafDesc


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


def feature_maps_format(feature_maps, inverse=False):
    """List of ``[B, M, C, h_l, w_l]`` maps <-> (``[B, M, sum(h_l*w_l), C]``, shapes, start offsets)
    — ``ops/deformable_aggregation.py:78-117``."""
    if not inverse:
        bs, num_cams, channels = feature_maps[0].shape[:3]
        shapes, starts, total = [], [], 0
        flat = []
        for fm in feature_maps:
            h, w = fm.shape[-2:]
            shapes.append((h, w))
            starts.append(total)
            total += h * w
            flat.append(fm.reshape(bs, num_cams, channels, h * w))
        col = torch.cat(flat, dim=-1).permute(0, 1, 3, 2)
        dev = col.device
        return [col, torch.tensor(shapes, dtype=torch.int64, device=dev),
                torch.tensor(starts, dtype=torch.int64, device=dev)]
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
