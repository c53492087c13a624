"""Import-name shim: `import local_aggregate` resolves to the B200-native op, so the reference's
`GaussianHead` (model/head/gaussian_head.py:30-39) constructs it unchanged."""
from gaussianformer_b200.splat import LocalAggregator as LocalAggregator  # noqa: F401

__all__ = ["LocalAggregator"]
