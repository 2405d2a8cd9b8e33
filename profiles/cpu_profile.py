"""Host-side (Python) cost of one bench step: cProfile over a few steps, top functions by
cumulative and by own time.  The step is GPU bound only while the host stays ahead; at the coarse
levels of the UNet the kernels take 10-30 us and the host path per layer decides the step time.

    python profiles/cpu_profile.py [--model MinkUNet34C] [--clouds 8] [--voxels 100000]"""
import argparse
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import minkowskiengine_b200 as ME  # noqa: E402
from bench import make_batch  # noqa: E402
from examples.minkunet import minkunet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clouds", type=int, default=8)
    ap.add_argument("--voxels", type=int, default=100000)
    ap.add_argument("--model", default="MinkUNet34C")
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = minkunet(a.model, ME, 3, 20, 3).to(dev)
    opt = torch.optim.SGD(net.parameters(), lr=1e-2)
    crit = torch.nn.CrossEntropyLoss()
    c, f, l = make_batch(a.clouds, a.voxels, 0)
    c, f, l = c.to(dev), f.to(dev).bfloat16(), l.to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        out = net(ME.SparseTensor(f, c))
        loss = crit(out.F.float(), l)
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    # host time to ENQUEUE a step (no sync inside except the per-map count reads)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"host enqueue time per step {1e3 * (t1 - t0) / a.steps:.2f} ms; "
          f"with the final sync {1e3 * (t2 - t0) / a.steps:.2f} ms")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(a.steps):
        step()
    pr.disable()
    torch.cuda.synchronize()
    for key in ("cumulative", "tottime"):
        buf = io.StringIO()
        pstats.Stats(pr, stream=buf).sort_stats(key).print_stats(28)
        txt = buf.getvalue()
        print(f"---- top by {key} ({a.steps} steps) ----")
        print("\n".join(ln[:150] for ln in txt.splitlines()[4:44]))


if __name__ == "__main__":
    main()
