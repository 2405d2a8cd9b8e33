"""Host-side groundwork for tile-staged convolution (minkowskiengine_b200/tile_order.py),
checked on the CPU against straightforward numpy: the permutation is a permutation that keeps
clouds apart and puts spatial neighbours into the same tiles; the staging tables reproduce the
neighbour table exactly (rows[tile][local] == nbr) with no duplicate in a tile's list."""
import numpy as np
import torch

from minkowskiengine_b200 import tile_order as TO
from oracle import oracle_np as O
from oracle.oracle_np import surface_cloud


def _table(coords_np, ks=3):
    offs = O.region_offsets(O.HYPER_CUBE, [ks] * 3, [1] * 3, [1] * 3)
    im, om = O.kernel_map(coords_np, coords_np, offs)
    K, n = len(offs), len(coords_np)
    nbr = np.full((K, n), -1, np.int32)
    for k in range(K):
        nbr[k, om[k]] = im[k]
    return torch.from_numpy(nbr)


def test_morton_order_groups_clouds_and_neighbours():
    c = torch.cat([surface_cloud(3000, 1, batch=1), surface_cloud(3000, 0, batch=0)])
    perm = TO.morton_order(c)
    assert sorted(perm.tolist()) == list(range(len(c)))
    b = c[perm, 0]
    assert bool((b[1:] >= b[:-1]).all())                       # batch-major
    nbr = _table(c.numpy())
    rows, counts, _ = TO.staging_tables(nbr, perm, tile=128)
    rows_id, counts_id, _ = TO.staging_tables(nbr, torch.arange(len(c)), tile=128)
    pairs = int((nbr >= 0).sum())
    # far fewer distinct source rows per tile than references: the reuse the kernels can exploit
    assert int(counts.sum()) < 0.25 * pairs and int(counts.sum()) < 0.3 * int(counts_id.sum())
    assert TO.morton_order(c[:0]).numel() == 0


def test_staging_tables_reproduce_the_neighbour_table():
    c = surface_cloud(5000, 3)
    nbr = _table(c.numpy())
    K, n = nbr.shape
    for perm in (torch.arange(n), TO.morton_order(c), torch.randperm(n, generator=torch.Generator().manual_seed(0))):
        for tile in (128, 64):
            rows, counts, local = TO.staging_tables(nbr, perm, tile=tile)
            n_tiles = (n + tile - 1) // tile
            assert rows.shape[0] == n_tiles and local.shape == (K, n_tiles * tile)
            r, cnt, loc = rows.numpy(), counts.numpy(), local.numpy()
            g = np.full((K, n_tiles * tile), -1, np.int64)
            g[:, :n] = nbr.numpy()[:, perm.numpy()]
            for t in range(n_tiles):
                lst = r[t, :cnt[t]]
                assert (lst >= 0).all() and (np.diff(lst) > 0).all()      # distinct, ascending
                assert (r[t, cnt[t]:] == -1).all()
                blk, lb = g[:, t * tile:(t + 1) * tile], loc[:, t * tile:(t + 1) * tile]
                assert ((lb == TO.MISSING) == (blk < 0)).all()
                hit = blk >= 0
                assert (lst[lb[hit]] == blk[hit]).all()
                assert set(lst.tolist()) == set(blk[hit].tolist())


def test_ordered_tables_permute_columns_only():
    c = surface_cloud(2000, 5)
    nbr = _table(c.numpy())
    tile_nbr, row_ids = TO.ordered_tables(nbr, c, 1)
    assert tile_nbr.dtype == torch.int32 and row_ids.dtype == torch.int32
    assert sorted(row_ids.tolist()) == list(range(len(c)))
    assert torch.equal(tile_nbr, nbr[:, row_ids.long()])
    # strided level: coordinates are multiples of the tensor stride
    c2 = c.clone()
    c2[:, 1:] = torch.div(c2[:, 1:], 2, rounding_mode="floor") * 2
    c2 = torch.unique(c2, dim=0)
    p2 = TO.morton_order(c2, 2)
    p1 = TO.morton_order(torch.cat([c2[:, :1], c2[:, 1:] // 2], 1), 1)
    assert torch.equal(p1, p2)
