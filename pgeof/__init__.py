"""Drop-in alias: ``import pgeof`` -> the B200-native implementation.

Same re-exports as the reference's ``src/pgeof/__init__.py:1-9``.
"""
from point_geometric_features_b200 import (  # noqa: F401
    EFeatureID,
    compute_features,
    compute_features_multiscale,
    compute_features_optimal,
    compute_features_selected,
    knn_search,
    pgeof_ext,
    radius_search,
)
from point_geometric_features_b200 import __version__  # noqa: F401
