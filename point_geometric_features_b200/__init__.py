"""B200-native implementation of the pgeof hot path (neighbour search + neighbourhood-PCA features).

Drop-in for ``pgeof`` (drprojects/point_geometric_features, ``src/pgeof/__init__.py:1-9``): the
same seven names with the same arguments, defaults and dtype rules, executed by hand-written
sm_100a CUDA kernels behind the C ABI of ``include/pgeof_b200.h``.  ``import pgeof`` resolves to
this package through the ``pgeof/`` shim at the repository root.

There is no CPU fallback: importing works anywhere (so argument checking can be tested without a
GPU), but every compute call raises ``RuntimeError`` when no CUDA device is present, and the
import itself fails loudly if the compiled extension is missing.
"""
import os as _os

_HERE = _os.path.dirname(_os.path.abspath(__file__))

try:
    from .pgeof_ext import (  # noqa: F401
        EFeatureID,
        compute_features,
        compute_features_multiscale,
        compute_features_optimal,
        compute_features_selected,
        knn_search,
        radius_search,
    )
    from . import pgeof_ext  # noqa: F401
except ImportError as _e:  # pragma: no cover - build problem, never a silent fallback
    raise ImportError(
        "point_geometric_features_b200: the compiled extension is missing or does not load (%s). "
        "Build it in-tree with `python build_native.py` "
        "(or `python -c 'import __graft_entry__ as g; g.build()'`)." % (_e,)
    ) from _e

# extensions beyond the reference API
from .pgeof_ext import (  # noqa: E402,F401
    device_count,
    get_eig_order,
    knn_features,
    knn_search_csr,
    launch_count,
    profile_enable,
    profile_read,
    profile_reset,
    radius_search_csr,
    reset_launch_count,
    set_eig_order,
    slab_select,
    trim,
)

LIBRARY_PATH = _os.path.join(_HERE, "libpgeof_b200.so")
HEADER_PATH = _os.path.join(_os.path.dirname(_HERE), "include", "pgeof_b200.h")

__all__ = [
    "EFeatureID", "compute_features", "compute_features_multiscale", "compute_features_optimal",
    "knn_search", "radius_search", "compute_features_selected",
]
__version__ = "0.3.3+b200.1"
