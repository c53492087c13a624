"""The C ABI called directly (ctypes, raw device pointers): status codes, native-boundary inputs
(points_int / means_int / radii supplied, cov as [G,6]) and workspace checks."""
import ctypes

import numpy as np
import pytest
import torch

import helpers as h
from gaussianformer_b200 import _lib
from gaussianformer_b200.splat import _make_desc, splat_forward_raw, splat_backward_raw

pytestmark = pytest.mark.gpu


def test_native_boundary_inputs_match_fused_prep():
    kw, inp, variant = h.splat_case("tiny", 8, True)
    a, pi, mi, radii, cov6, dims = h.oracle_prep(kw, inp, variant)
    dev = "cuda"
    T = lambda x, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(x)).to(dev, dt)
    N, G = a["pts"].shape[0], a["means"].shape[0]
    desc = _make_desc(G, N, 18, *dims, _lib.GF_SPLAT_BASE, 1, 6, kw["pc_min"], kw["grid_size"],
                      kw["scale_multiplier"], 0)
    (logits, _, _, _), ws = splat_forward_raw(desc, T(a["pts"]), T(a["means"]), T(a["opa"]), T(a["sem"]), T(cov6),
                                              points_int=T(pi, torch.int32), means_int=T(mi, torch.int32),
                                              radii=T(radii, torch.int32))
    logits = logits[0]                                  # leading batch dimension of the batched ABI (B = 1)
    ref = h.oracle_forward(kw, inp, variant)
    h.assert_close(logits.cpu().numpy(), ref["logits"], what="native-boundary logits")
    # and the module path (fused prep, cov 3x3) gives bit-identical output
    m = h.make_module(kw, variant)
    t = h.to_dev(inp)
    out = m(t["pts"], t["means"], t["opa"], t["sem"], t["scales"], t["cov"])
    assert torch.equal(out, logits)
    g = torch.randn(N, 18, generator=torch.Generator().manual_seed(3)).to(dev)
    gm, go, gs, gc, _, _ = (None if x is None else x[0] for x in splat_backward_raw(
        desc, T(a["pts"]), T(a["means"]), T(a["opa"]), T(a["sem"]), T(cov6), (g, None, None), (None, None, None),
        points_int=T(pi, torch.int32), means_int=T(mi, torch.int32), radii=T(radii, torch.int32)))
    rm, ro, rs, rc = h.oracle_backward(kw, inp, variant, (g.cpu().numpy(),))
    r32 = h.oracle_backward(kw, inp, variant, (g.cpu().numpy(),), precision="f32")
    for name, mine, r, rr in (("means", gm, rm, r32[0]), ("opa", go, ro, r32[1]), ("sem", gs, rs, r32[2]), ("cov", gc, rc, r32[3])):
        h.assert_grad_parity(mine.cpu().numpy(), r, rr, what="native grad " + name)


def test_status_codes():
    L = _lib.lib()
    d = _make_desc(10, 10, 18, 4, 4, 4, 0, 1, 6, (0, 0, 0), 1.0, 3.0, 0)
    assert L.gf_splat_forward(ctypes.byref(d), None, None, None, 0, None) == 1          # GF_ERR_INVALID_ARG
    assert b"NULL" in L.gf_last_error()
    d_bad = _make_desc(10, 10, 7, 4, 4, 4, 0, 1, 6, (0, 0, 0), 1.0, 3.0, 0)
    assert L.gf_splat_forward_workspace_bytes(ctypes.byref(d_bad)) == 0
    assert b"C=7" in L.gf_last_error()
    t = torch.zeros(64, device="cuda")
    ins = _lib.SplatInputs(*([ctypes.c_void_p(t.data_ptr())] * 9))
    outs = _lib.SplatOutputs(*([ctypes.c_void_p(t.data_ptr())] * 5))
    rc = L.gf_splat_forward(ctypes.byref(d), ctypes.byref(ins), ctypes.byref(outs), ctypes.c_void_p(t.data_ptr()), 16, None)
    assert rc == 2                                                                        # GF_ERR_WORKSPACE
    assert 18 in _lib.supported_classes()
    assert L.gf_abi_version() == 2


def test_zero_points_still_initialises_the_status_word():
    """N == 0: nothing is rendered, but the preparation kernels run, so gf_splat_read_flags reports the Gaussian
    error bits instead of uninitialised memory."""
    from gaussianformer_b200.splat import read_flags
    G = 40
    gen = torch.Generator().manual_seed(0)
    means = (torch.rand(G, 3, generator=gen) * 4).cuda()
    means[3, 0] = 1e3                                   # outside the 8x8x8 grid
    d = _make_desc(G, 0, 18, 8, 8, 8, _lib.GF_SPLAT_BASE, 1, 6, (0.0, 0.0, 0.0), 0.5, 3.0, 0)
    cov6 = torch.tensor([[4.0, 4.0, 4.0, 0.0, 0.0, 0.0]]).repeat(G, 1).cuda()
    scales = torch.full((G, 3), 0.3).cuda()
    (_lg, _, _, _), ws = splat_forward_raw(d, torch.zeros(0, 3).cuda(), means, torch.ones(G).cuda(),
                                           torch.rand(G, 18, generator=gen).cuda(), cov6, scales=scales)
    assert read_flags(ws, ws.device) & _lib.GF_FLAG_MEAN_OUT_OF_GRID
