"""Per-layer timing of the convolution kernels (CUDA events, kernel map cached) on the layer
shapes MinkUNet34C runs at the bench's batch (8 x 100k-voxel surface clouds).

    python profiles/layer_bench.py [--clouds 8] [--voxels 100000] [--dtype bf16] [--json out]
Reports per layer: rows, pairs P, ms fwd / dgrad / wgrad, achieved algorithmic TFLOP/s
(2*P*Cin*Cout per pass) and compulsory GB/s (SURVEY.md §8d formula)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import minkowskiengine_b200 as ME  # noqa: E402
from minkowskiengine_b200 import backend  # noqa: E402
from examples.synthetic import surface_cloud  # noqa: E402

LAYERS = [  # (name, tensor stride level, cin, cout, ks, stride, transposed)
    ("conv0 3->32 k5", 1, 3, 32, 5, 1, False),
    ("conv1 32->32 k2s2", 1, 32, 32, 2, 2, False),
    ("block1 32->32 k3", 2, 32, 32, 3, 1, False),
    ("block2 64->64 k3", 4, 64, 64, 3, 1, False),
    ("block3 128->128 k3", 8, 128, 128, 3, 1, False),
    ("block4 256->256 k3", 16, 256, 256, 3, 1, False),
    ("block5 384->256 k3", 8, 384, 256, 3, 1, False),
    ("block6 192->128 k3", 4, 192, 128, 3, 1, False),
    ("block7 128->96 k3", 2, 128, 96, 3, 1, False),
    ("block7 96->96 k3", 2, 96, 96, 3, 1, False),
    ("convtr7 96->96 k2s2T", 2, 96, 96, 2, 2, True),
    ("block8 128->96 k3", 1, 128, 96, 3, 1, False),
    ("block8 96->96 k3", 1, 96, 96, 3, 1, False),
]


def time_ms(fn, iters):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        fn()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clouds", type=int, default=8)
    ap.add_argument("--voxels", type=int, default=100000)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--json", default=None)
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[a.dtype]
    coords = torch.cat([surface_cloud(a.voxels, j, batch=j) for j in range(a.clouds)]).to(dev)
    x0 = ME.SparseTensor(torch.zeros(len(coords), 1, device=dev), coords)
    mgr = x0.coordinate_manager
    keys = {1: x0.coordinate_map_key}
    for lvl in (2, 4, 8, 16):
        keys[lvl] = mgr.stride(keys[lvl // 2], 2)
    rows = []
    esz = 4 if dt == torch.float32 else 2
    for name, lvl, cin, cout, ks, st, tr in LAYERS:
        if a.only and a.only not in name:
            continue
        in_key = keys[lvl]
        if tr:
            out_key = keys[lvl // 2]
        else:
            out_key = keys[lvl * st] if st > 1 else in_key
        km = mgr._manager._kernel_map(in_key, out_key, [ks] * 3, [st] * 3, [1] * 3,
                                      ME.RegionType.HYPER_CUBE, torch.IntTensor(), tr, False)
        P = int((km.out_nbr >= 0).sum().item())
        n_in, n_out, K = km.n_in, km.n_out, km.K
        feats = (torch.rand(n_in, cin, device=dev) - 0.5).to(dt)
        gout = (torch.rand(n_out, cout, device=dev) - 0.5).to(dt)
        w = ((torch.rand(K, cin, cout, device=dev) - 0.5) / (cin * K) ** 0.5)
        wl = w.to(dt)
        t_f = time_ms(lambda: backend._conv_forward(feats, w if dt != torch.float32 else wl, km), a.iters)   # fp32 master -> packed operands
        t_d = time_ms(lambda: backend._conv_backward(feats, gout, w, km, True, False), a.iters)
        t_w = time_ms(lambda: backend._conv_backward(feats, gout, w, km, False, True), a.iters)
        flops = 2.0 * P * cin * cout
        byt_f = n_in * cin * esz + n_out * cout * esz + K * cin * cout * esz + P * 8
        r = {"layer": name, "n_in": n_in, "n_out": n_out, "K": K, "pairs": P,
             "ms_fwd": t_f, "ms_dgrad": t_d, "ms_wgrad": t_w,
             "tflops_fwd": flops / t_f / 1e9, "tflops_dgrad": flops / t_d / 1e9,
             "tflops_wgrad": flops / t_w / 1e9, "gbs_fwd": byt_f / t_f / 1e6}
        rows.append(r)
        print(f"{name:24s} n_in={n_in:7d} n_out={n_out:7d} K={K:3d} P={P:9d} | fwd {t_f:7.3f} ms "
              f"{r['tflops_fwd']:7.1f} TF/s {r['gbs_fwd']:7.0f} GB/s | dgrad {t_d:7.3f} ms "
              f"{r['tflops_dgrad']:7.1f} | wgrad {t_w:7.3f} ms {r['tflops_wgrad']:7.1f}", flush=True)
    if a.json:
        json.dump(rows, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
