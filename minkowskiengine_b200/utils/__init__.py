"""Collation / init helpers the networks on the path use
(reference: MinkowskiEngine/utils/collation.py:30-188, utils/init.py)."""
import math

import numpy as np
import torch


def batched_coordinates(coords, dtype=torch.int32, device=None):
    """list of [N_i, D] coordinate arrays -> [sum N_i, D+1] with the batch index in column 0."""
    assert isinstance(coords, (list, tuple))
    out = []
    for b, c in enumerate(coords):
        c = torch.as_tensor(np.asarray(c) if not isinstance(c, torch.Tensor) else c)
        c = torch.floor(c).to(dtype) if c.is_floating_point() else c.to(dtype)
        out.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=dtype), c], dim=1))
    res = torch.cat(out, dim=0) if out else torch.zeros((0, 1), dtype=dtype)
    return res.to(device) if device is not None else res


def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
    bcoords = batched_coordinates(coords, dtype=dtype, device=device)
    bfeats = torch.cat([torch.as_tensor(f) for f in feats], dim=0)
    if device is not None:
        bfeats = bfeats.to(device)
    if labels is None:
        return bcoords, bfeats
    blabels = torch.cat([torch.as_tensor(l) for l in labels], dim=0)
    return bcoords, bfeats, (blabels.to(device) if device is not None else blabels)


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
