"""libmeb200's batch-norm passes against torch.nn.BatchNorm1d (fp32 reference of the same op):
forward, running statistics, input/weight/bias gradients, inference mode."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("n,C", [(5000, 32), (12345, 96), (777, 256), (3, 8), (40000, 64)])
def test_batchnorm_matches_torch(ME, cuda, dtype, tol, n, C):
    torch.manual_seed(n + C)
    x = (torch.randn(n, C) * 2.0 + torch.linspace(-3, 3, C)).to(cuda)
    coords = torch.cat([torch.zeros(n, 1, dtype=torch.int32),
                        torch.arange(n, dtype=torch.int32).unsqueeze(1).repeat(1, 3)], 1).to(cuda)
    ours = ME.MinkowskiBatchNorm(C).to(cuda)
    ref = torch.nn.BatchNorm1d(C).to(cuda)
    with torch.no_grad():
        ours.bn.weight.uniform_(0.5, 1.5); ours.bn.bias.uniform_(-1, 1)
        ref.weight.copy_(ours.bn.weight); ref.bias.copy_(ours.bn.bias)
    xq = x.to(dtype)
    st = ME.SparseTensor(xq.clone().requires_grad_(True), coords)
    y = ours(st)
    xr = xq.float().clone().requires_grad_(True)
    yr = ref(xr)
    scale = yr.abs().max().item()
    assert (y.F.float() - yr).abs().max().item() / scale < tol
    assert torch.allclose(ours.bn.running_mean, ref.running_mean, atol=1e-4, rtol=1e-4)
    assert torch.allclose(ours.bn.running_var, ref.running_var, atol=1e-4, rtol=1e-3)
    assert int(ours.bn.num_batches_tracked) == 1
    g = torch.randn(n, C, device=cuda)
    y.F.backward(g.to(dtype))
    yr.backward(g.to(dtype).float())
    gscale = xr.grad.abs().max().item()
    assert (st.F.grad.float() - xr.grad).abs().max().item() / gscale < tol
    assert (ours.bn.weight.grad - ref.weight.grad).abs().max().item() / ref.weight.grad.abs().max().item() < max(tol, 1e-4)
    assert (ours.bn.bias.grad - ref.bias.grad).abs().max().item() / ref.bias.grad.abs().max().item() < max(tol, 1e-4)
    # inference mode uses the running statistics
    ours.eval(); ref.eval()
    with torch.no_grad():
        ye = ours(ME.SparseTensor(xq, coords)).F.float()
        yre = ref(xq.float())
    assert (ye - yre).abs().max().item() / yre.abs().max().item() < tol
