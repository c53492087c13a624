"""Mirror of ``model/encoder/gaussian_encoder/ops/__init__.py:1`` in the reference."""
from .deformable_aggregation import DeformableAggregationFunction, feature_maps_format  # noqa: F401

__all__ = ["DeformableAggregationFunction", "feature_maps_format"]
