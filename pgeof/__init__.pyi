from point_geometric_features_b200.pgeof_ext import (
    EFeatureID as EFeatureID,
    compute_features as compute_features,
    compute_features_multiscale as compute_features_multiscale,
    compute_features_optimal as compute_features_optimal,
    compute_features_selected as compute_features_selected,
    knn_search as knn_search,
    radius_search as radius_search,
)
from point_geometric_features_b200 import pgeof_ext as pgeof_ext

__version__: str
