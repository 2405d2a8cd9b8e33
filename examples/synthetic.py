"""Synthetic input generators shared by bench.py, the profiling tools and the tests.
Neutral ground: neither the product package nor the oracle (bench.py's product arm must not
import anything from oracle/)."""


def surface_cloud(n, seed, batch=0):
    """The survey's seeded 'surface' generator: int32 [n, 4] unique rows."""
    import torch
    g = torch.Generator().manual_seed(seed)
    out = torch.empty(0, 3, dtype=torch.int32)
    r = 0.5 * (n / (4 * 3.14159)) ** 0.5 * 1.45
    while len(out) < n:
        v = torch.randn(4 * n, 3, generator=g)
        v = v / v.norm(dim=1, keepdim=True)
        c = torch.floor(v * r * (1 + 0.35 * torch.sin(3 * v[:, :1]))).int()
        out = torch.unique(torch.cat([out, c]), dim=0)
        r *= 1.05
    out = out[torch.randperm(len(out), generator=g)[:n]]
    return torch.cat([torch.full((n, 1), batch, dtype=torch.int32), out], 1)
