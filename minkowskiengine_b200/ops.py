"""Feature-level helpers used by the networks on the path: `cat`, `MinkowskiLinear`
(reference: MinkowskiOps.py:40-160)."""
import torch
from torch.nn import Module

from .sparse_tensor import (COORDINATE_KEY_DIFFERENT_ERROR, COORDINATE_MANAGER_DIFFERENT_ERROR,
                            SparseTensor)


class MinkowskiLinear(Module):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.linear = torch.nn.Linear(in_features, out_features, bias=bias)

    def forward(self, input):
        output = self.linear(input.F)
        return SparseTensor(output, coordinate_map_key=input.coordinate_map_key,
                            coordinate_manager=input.coordinate_manager)

    def __repr__(self):
        return (f"{self.__class__.__name__}(in_features={self.linear.in_features}, "
                f"out_features={self.linear.out_features}, bias={self.linear.bias is not None})")


def cat(*sparse_tensors):
    """Concatenate features of tensors living on the SAME coordinate map
    (reference: MinkowskiOps.py:70-110)."""
    assert len(sparse_tensors) > 1, f"Invalid number of inputs. The input must be at least two len(sparse_tensors) > 1"
    first = sparse_tensors[0]
    assert isinstance(first, SparseTensor), "Inputs must be SparseTensors"
    for s in sparse_tensors:
        assert isinstance(s, SparseTensor), "Inputs must be either SparseTensors or TensorFields."
        assert first.coordinate_manager == s.coordinate_manager, COORDINATE_MANAGER_DIFFERENT_ERROR
        assert first.coordinate_map_key == s.coordinate_map_key, \
            COORDINATE_KEY_DIFFERENT_ERROR + str(first.coordinate_map_key) + " != " + str(s.coordinate_map_key)
    feats = torch.cat([s.F for s in sparse_tensors], dim=1)
    return SparseTensor(feats, coordinate_map_key=first.coordinate_map_key,
                        coordinate_manager=first.coordinate_manager)
