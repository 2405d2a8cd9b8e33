"""Collation / init helpers the networks on the path use
(reference: MinkowskiEngine/utils/collation.py:30-188, utils/init.py)."""
import math

import numpy as np
import torch


def batched_coordinates(coords, dtype=torch.int32, device=None):
    """list of [N_i, D] coordinate arrays -> [sum N_i, D+1] with the batch index in column 0
    (reference: utils/collation.py:30-93).  Device tensors stay on their device."""
    assert isinstance(coords, (list, tuple))
    out = []
    for b, c in enumerate(coords):
        c = torch.as_tensor(np.asarray(c) if not isinstance(c, torch.Tensor) else c)
        if device is not None:
            c = c.to(device)
        c = torch.floor(c).to(dtype) if c.is_floating_point() else c.to(dtype)
        out.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=dtype, device=c.device), c],
                             dim=1))
    res = torch.cat(out, dim=0) if out else torch.zeros((0, 1), dtype=dtype)
    return res.to(device) if device is not None else res


def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
    """reference: utils/collation.py:96-188 (lists of per-sample arrays -> one batch)."""
    bcoords = batched_coordinates(coords, dtype=dtype, device=device)
    bfeats = torch.cat([torch.as_tensor(f) if device is None else torch.as_tensor(f).to(device)
                        for f in feats], dim=0)
    if labels is None:
        return bcoords, bfeats
    blabels = torch.cat([torch.as_tensor(l) if device is None else torch.as_tensor(l).to(device)
                         for l in labels], dim=0)
    return bcoords, bfeats, blabels


def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100,
                    return_index=False, return_inverse=False, return_maps_only=False,
                    quantization_size=None, device="cuda", label_mode="reference"):
    """Voxelisation on the GPU (reference: utils/quantization.py:136-333 + src/quantization.cpp):
    floor(coordinates / quantization_size) -> int32, deduplicated by the device hash map
    (meb200_insert_and_map: the FIRST point of every voxel is kept and voxels are numbered by
    first occurrence, as the reference CPU path does).  Same arguments and return values as the
    reference; unlike it, labels are resolved on the device too (quantization.cpp:140-196).
    label_mode="reference" (default) reproduces the reference bit for bit, INCLUDING its indexing
    slip: when the points of voxel u disagree it writes `ignore_label` to
    colabels[inverse_mapping[u]] — the voxel that contains point number u — instead of
    colabels[u] (quantization.cpp:188-191; `val.first` is a voxel id used as a point id).
    label_mode="consistent" applies the documented intent: a voxel keeps the label of its first
    point if all its points agree and gets `ignore_label` otherwise."""
    assert label_mode in ("reference", "consistent")
    from ..backend import CoordinateMapManagerGPU_c10
    assert isinstance(coordinates, (np.ndarray, torch.Tensor)), \
        "Coords must be either np.array or torch.Tensor."
    assert coordinates.ndim == 2, \
        "The coordinates must be a 2D matrix. The shape of the input is " + str(coordinates.shape)
    if return_inverse:
        assert return_index, "return_reverse must be set with return_index"
    dev = torch.device(device)
    assert dev.type == "cuda", "minkowskiengine_b200.utils.sparse_quantize runs on CUDA only"
    was_numpy = isinstance(coordinates, np.ndarray)
    c = torch.as_tensor(coordinates).to(dev)
    f = None if features is None else torch.as_tensor(features).to(dev)
    lab = None if labels is None else torch.as_tensor(labels).to(dev)
    if f is not None:
        assert f.ndim == 2 and f.shape[0] == c.shape[0]
    if lab is not None:
        assert lab.shape[0] == c.shape[0]
    if quantization_size is not None:
        if isinstance(quantization_size, (list, tuple, np.ndarray, torch.Tensor)):
            assert len(quantization_size) == c.shape[1], \
                "Quantization size and coordinates size mismatch."
            q = torch.as_tensor(np.asarray(quantization_size, dtype=np.float64) if not
                                isinstance(quantization_size, torch.Tensor) else quantization_size)
            c = torch.floor(c / q.to(dev).to(c.dtype if c.is_floating_point() else torch.float32))
        elif np.isscalar(quantization_size):
            c = torch.floor(c) if quantization_size == 1 else torch.floor(c / quantization_size)
        else:
            raise ValueError("Not supported type for quantization_size.")
    elif c.is_floating_point():
        c = torch.floor(c)
    discrete = c.int().contiguous()
    manager = CoordinateMapManagerGPU_c10()
    _, (unique_map, inverse_map) = manager.insert_and_map(
        discrete, [1] * (discrete.shape[1] - 1), "")
    if inverse_map.numel() == 0:            # no duplicates: identity (see insert_and_map)
        inverse_map = torch.arange(discrete.shape[0], dtype=torch.int64, device=dev)

    def back(t):
        return t.cpu().numpy() if was_numpy else t

    if lab is not None:
        first = lab[unique_map]
        differs = (lab != first[inverse_map]).to(torch.int32)
        mixed = torch.zeros(unique_map.shape[0], dtype=torch.int32, device=dev)
        mixed.index_add_(0, inverse_map, differs)
        # a voxel whose first label already is ignore_label never "conflicts" in the reference
        conflict = (mixed > 0) & (first != ignore_label)
        if label_mode == "consistent":
            colabels = torch.where(conflict, torch.full_like(first, ignore_label), first)
        else:
            colabels = first.clone()
            target = inverse_map[torch.nonzero(conflict).flatten()]   # voxel of point number u
            colabels[target] = ignore_label
        ret = [back(discrete[unique_map])]
        if f is not None:
            ret.append(back(f[unique_map]))
        ret.append(back(colabels))
        if return_index:
            ret.append(back(unique_map))
        if return_inverse:
            ret.append(back(inverse_map))
        return ret[0] if len(ret) == 1 else tuple(ret)
    if return_maps_only:
        return (back(unique_map), back(inverse_map)) if return_inverse else back(unique_map)
    ret = [back(discrete[unique_map])]
    if f is not None:
        ret.append(back(f[unique_map]))
    if return_index:
        ret.append(back(unique_map))
    if return_inverse:
        ret.append(back(inverse_map))
    return ret[0] if len(ret) == 1 else tuple(ret)


def kaiming_normal_(tensor, a=0, mode="fan_in", nonlinearity="leaky_relu"):
    """Kaiming init for [K, Cin, Cout] kernels (reference: utils/init.py:7-41):
    fan_in = K*Cin, fan_out = K*Cout."""
    if tensor.dim() == 3:
        fan_in, fan_out = tensor.size(0) * tensor.size(1), tensor.size(0) * tensor.size(2)
    else:
        # 2-D kernels ([Cin, Cout], kernel volume 1) are treated like torch Linear weights by
        # the reference (utils/init.py:12-14): fan_in = size(1), fan_out = size(0).  Kept so
        # that a seeded initialisation reproduces the reference's weights bit for bit.
        fan_in, fan_out = tensor.size(1), tensor.size(0)
    fan = fan_in if mode == "fan_in" else fan_out
    gain = torch.nn.init.calculate_gain(nonlinearity, a)
    std = gain / math.sqrt(fan)
    with torch.no_grad():
        return tensor.normal_(0, std)
