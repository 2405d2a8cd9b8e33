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


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("n,C,with_res,training", [(12345, 96, True, True), (5000, 32, False, True),
                                                   (777, 256, True, False), (4000, 64, False, False)])
def test_fused_bn_relu_matches_torch(ME, cuda, dtype, tol, n, C, with_res, training):
    """fused_bn_relu = relu(bn(x) [+ residual]) (the residual-block tail, reference
    modules/resnet_block.py:52-68) against torch's three separate ops, forward and backward —
    including the eval-mode (running statistics) path under autograd."""
    torch.manual_seed(n + C + with_res)
    x = (torch.randn(n, C) * 2.0 + torch.linspace(-3, 3, C)).to(cuda).to(dtype)
    res = (torch.randn(n, C) * 1.5).to(cuda).to(dtype) if with_res else None
    coords = torch.cat([torch.zeros(n, 1, dtype=torch.int32),
                        torch.arange(n, dtype=torch.int32).unsqueeze(1).repeat(1, 3)], 1).to(cuda)
    ours = ME.MinkowskiBatchNorm(C).to(cuda)
    ref = torch.nn.BatchNorm1d(C).to(cuda)
    with torch.no_grad():
        ours.bn.weight.uniform_(0.5, 1.5); ours.bn.bias.uniform_(-1, 1)
        ours.bn.running_mean.uniform_(-1, 1); ours.bn.running_var.uniform_(0.5, 2.0)
        ref.load_state_dict(ours.bn.state_dict())
    ours.train(training); ref.train(training)
    xs = ME.SparseTensor(x.clone().requires_grad_(True), coords)
    rs = None
    if with_res:
        rs = ME.SparseTensor(res.clone().requires_grad_(True), coordinate_map_key=xs.coordinate_map_key,
                             coordinate_manager=xs.coordinate_manager)
    y = ME.fused_bn_relu(ours, xs, rs)
    xr = x.float().clone().requires_grad_(True)
    rr = res.float().clone().requires_grad_(True) if with_res else None
    yr = ref(xr)
    if with_res:
        yr = yr + rr
    yr = torch.relu(yr)
    scale = yr.abs().max().item()
    assert (y.F.float() - yr).abs().max().item() / scale < tol
    if training:
        assert torch.allclose(ours.bn.running_mean, ref.running_mean, atol=1e-4, rtol=1e-4)
        assert torch.allclose(ours.bn.running_var, ref.running_var, atol=1e-4, rtol=1e-3)
    g = torch.randn(n, C, device=cuda).to(dtype)
    y.F.backward(g)
    yr.backward(g.float())
    gscale = xr.grad.abs().max().item()
    # a bf16 output can round a tiny positive pre-activation to exactly the sign the fp32
    # reference does not have; such elements carry |y| < 2^-8 |scale| and are excluded
    agree = ((y.F.float() > 0) == (yr > 0))
    assert float((~agree).float().mean()) < 2e-3
    assert ((xs.F.grad.float() - xr.grad).abs() * agree).max().item() / gscale < 4 * tol
    wg, wr = ours.bn.weight.grad, ref.weight.grad
    assert (wg - wr).abs().max().item() / wr.abs().max().item() < max(4 * tol, 1e-4)
    assert (ours.bn.bias.grad - ref.bias.grad).abs().max().item() / ref.bias.grad.abs().max().item() < max(4 * tol, 1e-4)
    if with_res:
        assert ((rs.F.grad.float() - rr.grad).abs() * agree).max().item() / rr.grad.abs().max().item() < 4 * tol


def test_reduction_workspace_is_left_clean(ME, cuda):
    """The launch-lean passes share one workspace per stream (csrc/batchnorm.cu BnTail): after any
    sequence of layers of different widths it must read all zeros, and results must not depend on
    what ran before."""
    from minkowskiengine_b200 import _lib, normalization as N
    torch.manual_seed(3)
    outs = []
    for rep in range(2):
        for n, C in [(4000, 256), (12345, 96), (50, 8), (30000, 32), (777, 2048)]:
            x = torch.randn(n, C, device=cuda).to(torch.bfloat16) if rep == 0 else outs[0][(n, C)][0]
            bn = ME.MinkowskiBatchNorm(C).to(cuda)
            coords = torch.cat([torch.zeros(n, 1, dtype=torch.int32),
                                torch.arange(n, dtype=torch.int32).unsqueeze(1).repeat(1, 3)], 1).to(cuda)
            st = ME.SparseTensor(x.clone().requires_grad_(True), coords)
            y = bn(st)
            y.F.backward(torch.ones_like(y.F))
            if rep == 0:
                outs.append({}) if not outs else None
                outs[0][(n, C)] = (x, y.F.detach().clone(), st.F.grad.clone())
            else:
                _, y0, g0 = outs[0][(n, C)]
                assert torch.equal(y.F, y0) or (y.F.float() - y0.float()).abs().max() < 6e-2
                assert (st.F.grad.float() - g0.float()).abs().max() < 6e-2
    torch.cuda.synchronize()
    assert N._WORKSPACES, "the launch-lean path did not run"
    for ws in N._WORKSPACES.values():
        assert int(ws.count_nonzero()) == 0
