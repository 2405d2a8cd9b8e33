"""Spatially coherent row tiles for the tensor-core convolution (round-2 groundwork).

The convolution kernels walk the rows of the stationary side in tiles of 128.  With rows in the
caller's order (random for a voxelised cloud) the ~8.5 neighbours of a row are spread over the
whole cloud, so every (tile, offset) stage gathers ~40 rows that no other stage of the tile
shares.  If a tile instead holds 128 rows that are close in space, its 27 offsets keep hitting
the same few hundred input rows: on the bench's surface clouds 1131 gathered rows per tile are
only 234 distinct ones (4.8x reuse; `profiles/r1_notes.md`).  This module computes, with plain
torch ops (device agnostic, so it is checked on the CPU against numpy):

  * `morton_order`   - a permutation that sorts rows by (batch, Morton code of the coordinates);
  * `staging_tables` - for every tile of the permuted order the list of DISTINCT source rows its
                        offsets reference and, per (offset, tile row), the position in that
                        list (0xFFFF = no neighbour) - what a kernel needs to fetch each source
                        row once per tile into shared memory and build the per-offset operand
                        from there.

  * `ordered_tables` / `conv_forward_ordered` - the cheap first step: permute the columns of
                        the neighbour table so that the kernel's tiles walk the rows in Morton
                        order and let the L1 cache serve the re-reads
                        (`meb200_conv_forward_ordered`, `k_conv_rs` EXP bit 3).

Nothing here is on the product path yet; it exists to be measured at the start of round 2
(`profiles/layer_bench.py --ordered`).
"""
import torch

MISSING = 0xFFFF


def _spread_bits(x, dims):
    """Insert dims-1 zero bits between the low 21 bits of x (int64)."""
    out = torch.zeros_like(x)
    for b in range(21):
        out |= ((x >> b) & 1) << (b * dims)
    return out


def morton_order(coords, tensor_stride=1):
    """coords: [N, D+1] integer rows (batch, x_1..x_D) -> int64 permutation [N]: rows sorted by
    batch index, then by the Morton (Z-order) code of the coordinates in units of the tensor
    stride, ties in original order (stable)."""
    c = coords.long()
    D = c.shape[1] - 1
    if c.shape[0] == 0:
        return torch.empty(0, dtype=torch.long, device=coords.device)
    ts = torch.as_tensor(tensor_stride, dtype=torch.long, device=c.device)
    ts = ts.expand(D) if ts.dim() == 0 else ts
    g = torch.div(c[:, 1:], ts, rounding_mode="floor")
    g = g - g.min(dim=0).values                       # non-negative grid coordinates
    assert int(g.max()) < (1 << 21), "coordinate range too large for a 63-bit Morton code"
    code = torch.zeros(c.shape[0], dtype=torch.long, device=c.device)
    for a in range(D):
        code |= _spread_bits(g[:, a], D) << a
    # batch-major: clouds do not interleave (a tile then reads one cloud's rows)
    order = torch.argsort(code, stable=True)
    return order[torch.argsort(c[order, 0], stable=True)]


def staging_tables(nbr, perm, tile=128):
    """nbr: [K, N] int32 neighbour table (-1 = missing) of the rows being tiled; perm: [N]
    permutation (tile t holds rows perm[t*tile : (t+1)*tile]).

    Returns (rows, counts, local):
      rows   [n_tiles, U] int32 - distinct source rows referenced by tile t, ascending, padded
                                  with -1 (U = the largest count over the tiles);
      counts [n_tiles]    int32 - number of valid entries of rows[t];
      local  [K, n_tiles * tile] uint16 stored as int32 - position of nbr[k, perm[j]] in
                                  rows[j // tile], MISSING where there is no neighbour (also for
                                  the padding rows of the last tile)."""
    K, N = nbr.shape
    n_tiles = (N + tile - 1) // tile
    dev = nbr.device
    pad = n_tiles * tile - N
    g = nbr[:, perm.long()].long()
    if pad:
        g = torch.cat([g, torch.full((K, pad), -1, dtype=torch.long, device=dev)], 1)
    # [n_tiles, K * tile]: all references of a tile in one row
    flat = g.view(K, n_tiles, tile).permute(1, 0, 2).reshape(n_tiles, K * tile)
    vals, idx = flat.sort(dim=1)
    first = torch.ones_like(vals, dtype=torch.bool)
    first[:, 1:] = vals[:, 1:] != vals[:, :-1]
    valid = vals >= 0
    # rank among the DISTINCT VALID values of the tile
    rank = torch.cumsum((first & valid).long(), dim=1) - 1
    counts = (first & valid).sum(dim=1)
    U = max(int(counts.max()) if n_tiles else 0, 1)
    assert U <= MISSING, "too many distinct rows in one tile for 16-bit positions"
    rows = torch.full((n_tiles, U), -1, dtype=torch.int32, device=dev)
    t_idx = torch.arange(n_tiles, device=dev).unsqueeze(1).expand_as(vals)
    sel = first & valid
    rows[t_idx[sel], rank[sel]] = vals[sel].int()
    loc_sorted = torch.where(valid, rank, torch.full_like(rank, MISSING))
    loc = torch.empty_like(loc_sorted)
    loc.scatter_(1, idx, loc_sorted)
    local = loc.view(n_tiles, K, tile).permute(1, 0, 2).reshape(K, n_tiles * tile).int()
    return rows, counts.int(), local


def ordered_tables(nbr, coords, tensor_stride=1):
    """-> (tile_nbr [K, N] int32 with tile_nbr[k, j] = nbr[k, row_ids[j]], row_ids [N] int32):
    the arguments of `meb200_conv_forward_ordered` for the rows whose coordinates are `coords`."""
    perm = morton_order(coords, tensor_stride)
    return nbr[:, perm].contiguous(), perm.int().contiguous()


def conv_forward_ordered(in_feat, kernel, tile_nbr, row_ids, out_dtype=None):
    """Forward convolution with spatially ordered tiles (CUDA tensors, tensor-core shapes only);
    same result as backend._conv_forward on the unpermuted table."""
    from . import _lib, backend
    lib = _lib.load()
    code = _lib.dtype_code(in_feat.dtype)
    kernel = kernel.to(in_feat.dtype).contiguous()
    K, c_in, c_out = kernel.shape
    n_in, n_out = in_feat.shape[0], tile_nbr.shape[1]
    out = torch.empty((n_out, c_out), dtype=out_dtype or in_feat.dtype, device=in_feat.device)
    ws, ws_bytes = backend._workspace(n_in, n_out, c_in, c_out, K, code, in_feat.device)
    _lib.check(lib.meb200_conv_forward_ordered(
        _lib.ptr(in_feat), code, n_in, c_in, _lib.ptr(kernel), K, c_out, _lib.ptr(tile_nbr),
        _lib.ptr(row_ids), n_out, _lib.ptr(out), _lib.dtype_code(out.dtype), _lib.ptr(ws),
        ws_bytes, _lib.current_stream()))
    return out
