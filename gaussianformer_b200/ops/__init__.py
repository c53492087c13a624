"""Mirror of ``model/encoder/gaussian_encoder/ops/__init__.py:1`` in the reference, plus the opt-in fused
caller path (``deformable_aggregation_fused``)."""
from .deformable_aggregation import (DeformableAggregationFunction, DeformableAggregationFusedFunction,  # noqa: F401
                                     deformable_aggregation_fused, feature_maps_format, fused_supported)

__all__ = ["DeformableAggregationFunction", "feature_maps_format", "DeformableAggregationFusedFunction",
           "deformable_aggregation_fused", "fused_supported"]
