"""Type stub of the compiled module (the reference generates ``pgeof_ext.pyi`` + ``py.typed`` with nanobind's
stubgen, CMakeLists.txt:34-41; this one is written by hand for the pybind11 build).

Array arguments are numpy arrays (host flavour: results are numpy arrays) or CUDA tensors exposing ``__dlpack__`` /
``__cuda_array_interface__`` such as ``torch.Tensor`` (device flavour: results are torch tensors on the same device).
Dtypes are strict, as in pgeof: float32 ``(n, 3)`` clouds, uint32 1-D ``nn`` / ``nn_ptr`` (``nn_ptr`` may also be
uint64 / int64 -- an extension), no implicit conversion (``TypeError``)."""
import enum
from typing import Any, Sequence, Tuple, Union

Array = Any   # numpy.ndarray | torch.Tensor (CUDA) | any __dlpack__ / __cuda_array_interface__ exporter

abi_version: int


class EFeatureID(enum.IntEnum):
    Linearity = 0
    Planarity = 1
    Scattering = 2
    VerticalityPGEOF = 3
    Normal_x = 4
    Normal_y = 5
    Normal_z = 6
    Length = 7
    Surface = 8
    Volume = 9
    Curvature = 10
    K_optimal = 11
    Verticality = 12
    Eigentropy = 13


Linearity: EFeatureID
Planarity: EFeatureID
Scattering: EFeatureID
VerticalityPGEOF: EFeatureID
Normal_x: EFeatureID
Normal_y: EFeatureID
Normal_z: EFeatureID
Length: EFeatureID
Surface: EFeatureID
Volume: EFeatureID
Curvature: EFeatureID
K_optimal: EFeatureID
Verticality: EFeatureID
Eigentropy: EFeatureID

# ---- the reference surface (src/pgeof_ext.cpp:34-177) ---------------------------------------------------------------
def compute_features(xyz: Array, nn: Array, nn_ptr: Array, k_min: int = 1, verbose: bool = False) -> Array:
    """float32 (num_points, 11) in EFeatureID order; rows shorter than k_min stay 0."""

def compute_features_multiscale(xyz: Array, nn: Array, nn_ptr: Array, k_scales: Sequence[int], verbose: bool = False) -> Array:
    """float32 (num_points, n_scales, 11); scale s uses the first k_scales[s] entries of every row."""

def compute_features_optimal(xyz: Array, nn: Array, nn_ptr: Array, k_min: int = 1, k_step: int = 1, k_min_search: int = 1,
                             verbose: bool = False) -> Array:
    """float32 (num_points, 12); column 11 is the neighbourhood size of lowest eigenentropy."""

def knn_search(data: Array, query: Array, knn: int) -> Tuple[Array, Array]:
    """(uint32 indices (n, knn), float32 squared distances (n, knn)), ascending by (distance, index)."""

def radius_search(data: Array, query: Array, search_radius: float, max_knn: int) -> Tuple[Array, Array]:
    """(int32 indices padded with -1, float32 squared distances padded with 0), the max_knn nearest with d2 < r*r."""

def compute_features_selected(xyz: Array, search_radius: float, max_knn: int,
                              selected_features: Sequence[Union[EFeatureID, int]]) -> Array:
    """(num_points, n_features) in the dtype of xyz (float32 or float64)."""

# ---- extensions ------------------------------------------------------------------------------------------------------
def radius_search_csr(data: Array, query: Array, search_radius: float, max_knn: int) -> Tuple[Array, Array]:
    """(nn uint32, nn_ptr uint32): the CSR the README glue builds from radius_search, from one search."""

def knn_search_csr(data: Array, query: Array, knn: int, wide_offsets: bool = False) -> Tuple[Array, Array]:
    """(nn uint32, nn_ptr uint32 | 64-bit): kNN as CSR, no distances; 64-bit offsets beyond 2^32-1 neighbours."""

def knn_features(xyz: Array, knn: int, k_min: int = 1, return_neighbors: bool = False) -> Any:
    """knn_search(xyz, xyz, knn) + compute_features without materialising the neighbour lists."""

def slab_select(xyz: Array, rank: int, world: int, axis: int = 2) -> Tuple[Array, Array]:
    """(rows int64, query float32 (m, 3)) of spatial query shard `rank` of `world` (CUDA tensors)."""

def set_eig_order(order: str) -> None: ...
def get_eig_order() -> str: ...
def device_count() -> int: ...
def launch_count() -> int: ...
def reset_launch_count() -> None: ...
def trim() -> None: ...
def profile_enable(on: bool) -> None: ...
def profile_reset() -> None: ...
def profile_read(name: str) -> Tuple[float, int]: ...
