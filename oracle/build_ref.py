"""ORACLE — TEST INFRASTRUCTURE ONLY.

Compiles the UNMODIFIED reference CUDA ops from where their sources lie under /root/reference
into ``oracle/_ref/`` (git-ignored, but shipped to the GPU box with the snapshot) as four torch
extension modules, for sm_100a.  No reference source is copied into the repo: nvcc/g++ read the
files in place and only objects / .so files are written under ``oracle/_ref``.

    gf_ref_localagg            <- model/head/localagg            (local_aggregate._C)
    gf_ref_localagg_prob       <- model/head/localagg_prob       (local_aggregate_prob._C)
    gf_ref_localagg_prob_fast  <- model/head/localagg_prob_fast  (local_aggregate_prob_fast._C)
    gf_ref_daf                 <- model/encoder/gaussian_encoder/ops (deformable_aggregation_ext)

Used (a) on the B200 box to produce the golden fixtures in tests/golden (make_golden_ref.py),
(b) by GPU parity tests when present, (c) by bench.py as the "reference CUDA op on the same GPU"
side figure.  The reference's own build system (setup.py) is not run.

``load_ref(name)`` imports a built module (no compilation, no /root/reference needed).
"""
from __future__ import annotations

import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("GF_REFERENCE_ROOT", "/root/reference")

_SPLAT_SOURCES = ["src/aggregator_impl.cu", "src/forward.cu", "src/backward.cu", "local_aggregate.cu", "ext.cpp"]
MODULES = {
    "gf_ref_localagg": ("model/head/localagg", _SPLAT_SOURCES),
    "gf_ref_localagg_prob": ("model/head/localagg_prob", _SPLAT_SOURCES),
    "gf_ref_localagg_prob_fast": ("model/head/localagg_prob_fast", _SPLAT_SOURCES),
    "gf_ref_daf": ("model/encoder/gaussian_encoder/ops",
                   ["src/deformable_aggregation.cpp", "src/deformable_aggregation_cuda.cu"]),
}


def so_path(name: str) -> str:
    return os.path.join(OUT, name + ".so")


def build_all(verbose: bool = False) -> list[str]:
    """Compile every reference op that is not built yet.  Needs /root/reference (this container)."""
    if not os.path.isdir(REF):
        return []
    # the image's CC/CXX point at a wrapper that cannot link; torch's builder honours these
    os.environ["CC"] = "/usr/bin/gcc"
    os.environ["CXX"] = "/usr/bin/g++"
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils.cpp_extension import load

    built = []
    for name, (subdir, sources) in MODULES.items():
        if os.path.exists(so_path(name)):
            built.append(name)
            continue
        bdir = os.path.join(OUT, "build_" + name)
        os.makedirs(bdir, exist_ok=True)
        srcs = [os.path.join(REF, subdir, s) for s in sources]
        load(name=name, sources=srcs, build_directory=bdir, verbose=verbose, is_python_module=False,
             extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fno-gnu-unique"],
             extra_include_paths=[os.path.join(REF, subdir)])
        os.replace(os.path.join(bdir, name + ".so"), so_path(name))
        built.append(name)
    return built


def available(name: str) -> bool:
    return os.path.exists(so_path(name))


def load_ref(name: str):
    """Import a prebuilt reference module (pybind11, needs torch imported first)."""
    import torch  # noqa: F401  (registers the torch shared libraries the module links against)

    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, so_path(name))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[name] = mod
    return mod


if __name__ == "__main__":
    print(build_all(verbose="-v" in sys.argv))
