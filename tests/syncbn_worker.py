"""Worker for tests/test_gpu_zz_syncbn.py: run under torch.distributed.run (or alone with
WORLD_SIZE unset = one rank).  Every rank builds the same [n, C] matrix, keeps its slice of the
rows, runs MinkowskiSyncBatchNorm forward + backward on the slice and checks the result against
torch.nn.BatchNorm1d over the FULL matrix (what synchronised statistics must reproduce)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import minkowskiengine_b200 as ME  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29531")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        for dtype, tol in ((torch.float32, 5e-5), (torch.bfloat16, 2e-2)):
            for n, C in ((6001, 32), (20000, 96)):
                torch.manual_seed(n + C)
                x = (torch.randn(n, C) * 2.0 + torch.linspace(-3, 3, C)).to(dev).to(dtype)
                g = torch.randn(n, C).to(dev).to(dtype)
                lo, hi = rank * n // world, (rank + 1) * n // world     # ragged split
                net = ME.MinkowskiSyncBatchNorm.convert_sync_batchnorm(
                    torch.nn.Sequential(ME.MinkowskiBatchNorm(C))).to(dev)
                ref = torch.nn.BatchNorm1d(C).to(dev)
                with torch.no_grad():
                    net[0].bn.weight.uniform_(0.5, 1.5)
                    net[0].bn.bias.uniform_(-1, 1)
                    dist.broadcast(net[0].bn.weight, 0)
                    dist.broadcast(net[0].bn.bias, 0)
                    ref.weight.copy_(net[0].bn.weight)
                    ref.bias.copy_(net[0].bn.bias)
                coords = torch.cat([torch.zeros(hi - lo, 1, dtype=torch.int32),
                                    torch.arange(lo, hi, dtype=torch.int32).unsqueeze(1).repeat(1, 3)],
                                   1).to(dev)
                xs = x[lo:hi].clone().requires_grad_(True)
                if world > 1:
                    yf = net(ME.SparseTensor(xs, coords)).F
                else:   # the module skips the exchange for one rank; drive the sync path directly
                    from minkowskiengine_b200 import normalization
                    yf = normalization._batch_norm(net[0].bn, xs, dist.group.WORLD)
                yf.backward(g[lo:hi])
                xr = x.float().clone().requires_grad_(True)
                yr = ref(xr)
                yr.backward(g.float())
                e_y = (yf.float() - yr[lo:hi]).abs().max().item() / yr.abs().max().item()
                e_dx = (xs.grad.float() - xr.grad[lo:hi]).abs().max().item() / xr.grad.abs().max().item()
                gw, gb = net[0].bn.weight.grad.clone(), net[0].bn.bias.grad.clone()
                dist.all_reduce(gw)
                dist.all_reduce(gb)
                e_w = (gw - ref.weight.grad).abs().max().item() / ref.weight.grad.abs().max().item()
                e_b = (gb - ref.bias.grad).abs().max().item() / ref.bias.grad.abs().max().item()
                e_rm = (net[0].bn.running_mean - ref.running_mean).abs().max().item()
                e_rv = ((net[0].bn.running_var - ref.running_var).abs() / ref.running_var).max().item()
                ok = e_y < tol and e_dx < tol and e_w < max(tol, 1e-4) and e_b < max(tol, 1e-4) \
                    and e_rm < 1e-4 and e_rv < 1e-3
                print(f"rank {rank} {dtype} n={n} C={C}: y {e_y:.2e} dx {e_dx:.2e} dw {e_w:.2e} "
                      f"db {e_b:.2e} rm {e_rm:.2e} rv {e_rv:.2e} {'ok' if ok else 'FAIL'}", flush=True)
                if not ok:
                    sys.exit(1)
        print(f"rank {rank}: SYNCBN_OK", flush=True)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
