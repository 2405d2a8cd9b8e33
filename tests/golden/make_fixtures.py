"""Generates tests/golden/ref_*.npz by running the COMPILED REFERENCE (oracle/_ref, built from
/root/reference by oracle/build_ref.py) on seeded inputs.  Run where the reference is available:

    python tests/golden/make_fixtures.py

The fixtures pin the numpy oracle (tests/test_oracle.py, CPU) and the CUDA path
(tests/test_gpu_golden.py) to outputs of the reference itself.  Derived-map row order is
implementation-defined in the reference, so coordinates are stored in lexicographic order with
features permuted accordingly and kernel maps as sorted (k, in coord, out coord) triples."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_np as O  # noqa: E402
from oracle import ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def cloud(n, extent, seed, D=3, batches=1):
    g = torch.Generator().manual_seed(seed)
    c = torch.randint(-extent, extent, (n, D), generator=g, dtype=torch.int32)
    b = torch.randint(0, batches, (n, 1), generator=g, dtype=torch.int32)
    return torch.cat([b, c], 1)


def kmap_triples(ME, x, y, K, **kw):
    kd = x.coordinate_manager.kernel_map(x.coordinate_map_key, y.coordinate_map_key, **kw)
    im = [kd[k][0].numpy().astype(np.int64) if k in kd else np.zeros(0, np.int64) for k in range(K)]
    om = [kd[k][1].numpy().astype(np.int64) if k in kd else np.zeros(0, np.int64) for k in range(K)]
    return O.kernel_map_triples(x.C.numpy(), y.C.numpy(), im, om)


def conv_case(ME, name, D, n, extent, cin, cout, ks, stride, seed, transpose_back=False):
    torch.manual_seed(seed)
    coords = cloud(n, extent, seed, D, batches=2)
    feats = torch.rand(n, cin)
    x = ME.SparseTensor(feats, coords, requires_grad=True)        # dedups (first occurrence wins)
    conv = ME.MinkowskiConvolution(cin, cout, kernel_size=ks, stride=stride, dimension=D)
    y = conv(x)
    gout = torch.rand(y.F.shape)
    y.F.backward(gout)
    out = {"in_coords": x.C.numpy(), "in_feats": x.F.detach().numpy(), "raw_coords": coords.numpy(),
           "raw_feats": feats.numpy(), "weight": conv.kernel.detach().numpy()}
    oc, of, perm = O.canonical_rows(y.C.numpy(), y.F.detach().numpy())
    out.update(out_coords=oc, out_feats=of, grad_out=gout.numpy()[perm],
               grad_in=x.F.grad.numpy(), grad_weight=conv.kernel.grad.numpy(),
               kmap=kmap_triples(ME, x, y, ks ** D, stride=stride, kernel_size=ks),
               meta=np.array([D, cin, cout, ks, stride]))
    np.savez_compressed(os.path.join(HERE, f"ref_conv_{name}.npz"), **out)
    print("wrote", name, len(x), "->", len(y), "pairs", len(out["kmap"]))


def pool_case(ME, name, mode, ks, stride, seed):
    D, C, n = 3, 5, 1500
    torch.manual_seed(seed)
    coords = cloud(n, 12, seed, D, 2)
    feats = torch.randn(n, C)
    x = ME.SparseTensor(feats, coords, requires_grad=True)
    layer = {"avg": ME.MinkowskiAvgPooling, "sum": ME.MinkowskiSumPooling,
             "max": ME.MinkowskiMaxPooling}[mode](kernel_size=ks, stride=stride, dimension=D)
    y = layer(x)
    gout = torch.rand(y.F.shape)
    y.F.backward(gout)
    oc, of, perm = O.canonical_rows(y.C.numpy(), y.F.detach().numpy())
    np.savez_compressed(os.path.join(HERE, f"ref_pool_{name}.npz"), in_coords=x.C.numpy(),
                        in_feats=x.F.detach().numpy(), out_coords=oc, out_feats=of,
                        grad_out=gout.numpy()[perm], grad_in=x.F.grad.numpy(),
                        meta=np.array([ks, stride, {"sum": 0, "avg": 1, "max": 2}[mode]]))
    print("wrote pool", name)


def transpose_case(ME, seed):
    D, cin, cmid, cout = 3, 4, 6, 3
    torch.manual_seed(seed)
    coords = cloud(1200, 10, seed, D, 2)
    feats = torch.rand(1200, cin)
    x = ME.SparseTensor(feats, coords, requires_grad=True)
    down = ME.MinkowskiConvolution(cin, cmid, kernel_size=2, stride=2, dimension=D)
    up = ME.MinkowskiConvolutionTranspose(cmid, cout, kernel_size=2, stride=2, dimension=D)
    z = up(down(x))
    assert z.coordinate_map_key == x.coordinate_map_key
    gout = torch.rand(z.F.shape)
    z.F.backward(gout)
    np.savez_compressed(os.path.join(HERE, "ref_conv_transpose_pair.npz"), in_coords=x.C.numpy(),
                        in_feats=x.F.detach().numpy(), w_down=down.kernel.detach().numpy(),
                        w_up=up.kernel.detach().numpy(), out_feats=z.F.detach().numpy(),
                        grad_out=gout.numpy(), grad_in=x.F.grad.numpy(),
                        grad_w_down=down.kernel.grad.numpy(), grad_w_up=up.kernel.grad.numpy())
    print("wrote transpose pair")


def main():
    ME = ref.import_reference()
    conv_case(ME, "cfg0_k3s1", 3, 1000, 16, 16, 16, 3, 1, seed=0)     # BASELINE configs[0] shape
    conv_case(ME, "k3s2", 3, 2000, 12, 8, 12, 3, 2, seed=1)
    conv_case(ME, "k2s2", 3, 2000, 12, 5, 7, 2, 2, seed=2)
    conv_case(ME, "4d_k3", 4, 2500, 5, 4, 6, 3, 1, seed=3)
    conv_case(ME, "k5s1_cin3", 3, 800, 8, 3, 8, 5, 1, seed=4)
    pool_case(ME, "max_k2s2", "max", 2, 2, 5)
    pool_case(ME, "avg_k3s2", "avg", 3, 2, 6)
    pool_case(ME, "sum_k2s2", "sum", 2, 2, 7)
    transpose_case(ME, 8)


if __name__ == "__main__":
    main()
