"""Host-side checks of the torch adaptor (argument handling only: every search needs a GPU)."""
import pytest
import torch

from point_geometric_features_b200 import torch_adaptor as ta


def test_segments_of_a_sorted_batch_vector():
    assert ta._segments(None, 7) == [(0, 7)]
    assert ta._segments(torch.tensor([0, 0, 1, 1, 1, 3]), 6) == [(0, 2), (2, 5), (5, 5), (5, 6)]
    with pytest.raises(ValueError):
        ta._segments(torch.tensor([1, 0]), 2)
    with pytest.raises(ValueError):
        ta._segments(torch.tensor([0]), 2)


def test_frnn_signature_validation():
    p = torch.zeros(2, 5, 3)
    with pytest.raises(ValueError):
        ta.frnn_grid_points(p, torch.zeros(3, 5, 3), K=2, r=1.0)
    with pytest.raises(ValueError):
        ta.frnn_grid_points(p[0], p[0], K=2, r=1.0)
    with pytest.raises(ValueError):
        ta.frnn_grid_points(p, p)                      # K and r are required, as in FRNN
    with pytest.raises(ValueError):
        ta.knn_2(p[0], p[0], 2, batch_search=torch.zeros(5, dtype=torch.int64))
