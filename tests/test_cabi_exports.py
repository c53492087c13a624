"""CPU-only: the C-ABI library loads and exports every symbol include/gf_b200.h declares;
the ctypes structs match the header's field order; the op modules fail loudly on CPU tensors."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions(name="gf_b200.h"):
    src = open(os.path.join(ROOT, "include", name)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from gaussianformer_b200 import _lib
    from gaussianformer_b200.csrc import build
    build.build()
    L = _lib.lib()
    names = _header_functions()
    assert set(names) == set(_lib.EXPORTS), (names, _lib.EXPORTS)
    debug = _header_functions("gf_b200_debug.h")          # measurement hooks live in their own header
    assert set(debug) == set(_lib.DEBUG_EXPORTS), (debug, _lib.DEBUG_EXPORTS)
    for n in names + debug:
        assert hasattr(L, n), n
    assert L.gf_abi_version() == 2


def test_struct_layouts_follow_the_header():
    from gaussianformer_b200 import _lib
    src = open(os.path.join(ROOT, "include", "gf_b200.h")).read()
    for cname, cls in (("gf_splat_desc", _lib.SplatDesc), ("gf_splat_inputs", _lib.SplatInputs),
                       ("gf_splat_outputs", _lib.SplatOutputs), ("gf_splat_grads", _lib.SplatGrads),
                       ("gf_daf_desc", _lib.DafDesc),
                       ("gf_daf_format_desc", _lib.DafFormatDesc)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), src, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.replace("*", " ").split()[-1:] if "," not in decl else None
            if names is None:   # "int32_t H, W, D"
                first, *rest = decl.split(",")
                names = [first.replace("*", " ").split()[-1]] + [r.strip().lstrip("*") for r in rest]
            fields += [n.split("[")[0] for n in names]
        assert fields == [f[0] for f in cls._fields_], (cname, fields)


def test_ops_refuse_cpu_tensors():
    from gaussianformer_b200.ops import DeformableAggregationFunction as DAF
    from gaussianformer_b200.splat import LocalAggregator
    m = LocalAggregator(3, 8, 8, 4, [0.0, 0.0, 0.0], 0.5)
    assert "pc_min" in dict(m.named_buffers()) and m.pc_min.shape == (1, 3)
    z = torch.zeros(1, 4, 3)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        m(z, z, torch.zeros(1, 4), torch.zeros(1, 4, 18), z, torch.zeros(1, 4, 3, 3))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        DAF.apply(torch.zeros(1, 1, 4, 8), torch.tensor([[2, 2]]), torch.tensor([0]), torch.zeros(1, 3, 1, 2),
                  torch.zeros(1, 3, 1, 1, 2))


def test_reference_import_names_resolve():
    import local_aggregate
    import local_aggregate_prob
    import local_aggregate_prob_fast
    from gaussianformer_b200 import splat
    assert local_aggregate.LocalAggregator is splat.LocalAggregator
    assert local_aggregate_prob.LocalAggregator is splat.LocalAggregatorProb
    assert local_aggregate_prob_fast.LocalAggregator is splat.LocalAggregatorProbFast
    # constructor kwargs of the shipped configs (config/nuscenes_gs25600_solid.py:185-190,
    # config/prob/nuscenes_gs6400.py:245-250)
    cuda_kwargs = dict(scale_multiplier=3, H=200, W=200, D=16, pc_min=[-50.0, -50.0, -5.0], grid_size=0.5)
    local_aggregate.LocalAggregator(**cuda_kwargs)
    local_aggregate_prob.LocalAggregator(**dict(cuda_kwargs, scale_multiplier=4))
    local_aggregate_prob_fast.LocalAggregator(**dict(cuda_kwargs, scale_multiplier=4))
