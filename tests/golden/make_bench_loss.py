"""Generates tests/golden/bench_step0_loss.json: the step-0 loss of bench.py's workloads computed
by the COMPILED REFERENCE (oracle/_ref: the reference's own fp32 CPU path) on bench.py's exact
synthetic batch and seed-0 initial weights.  bench.py asserts its own step-0 loss (bf16, GPU)
against these values.  Run where the reference is available:

    python tests/golden/make_bench_loss.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bench import make_batch, weights_digest  # noqa: E402
from examples.minkunet import minkunet  # noqa: E402
from oracle import ref  # noqa: E402


def main():
    REF = ref.import_reference()
    torch.set_num_threads(min(os.cpu_count() or 1, 16))
    out = {"torch": torch.__version__, "cases": {}}
    for tag, model, clouds, voxels in (("cfg3", "MinkUNet34C", 8, 100_000),
                                       ("cfg2", "MinkUNet14", 1, 50_000)):
        torch.manual_seed(0)
        net = minkunet(model, REF, 3, 20, 3)
        coords, feats, labels = make_batch(clouds, voxels, seed0=0)
        feats = feats.bfloat16().float()     # bench feeds bf16 features
        loss = torch.nn.functional.cross_entropy(net(REF.SparseTensor(feats, coords)).F, labels)
        out["cases"][tag] = {"model": model, "clouds": clouds, "voxels": voxels,
                             "loss": float(loss), "weights_digest": weights_digest(net)}
        print(tag, out["cases"][tag], flush=True)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_step0_loss.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
