"""Network-level parity on the GPU box: MinkUNet through this package (CUDA) against the
compiled reference CPU path (oracle/_ref) with identical weights and inputs."""
import numpy as np
import pytest
import torch

from examples.minkunet import minkunet
from oracle import oracle_np as O
from oracle import ref

pytestmark = pytest.mark.gpu


def _run(MEh, net, coords, feats, dev):
    x = MEh.SparseTensor(feats.to(dev), coords.to(dev))
    out = net(x)
    loss = out.F.float().pow(2).mean()
    loss.backward()
    return out, float(loss)


@pytest.mark.parametrize("name,n", [("MinkUNet14", 6000), ("MinkUNet34C", 4000)])
def test_minkunet_fp32_matches_reference(ME, cuda, name, n):
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    REF = ref.import_reference()
    torch.manual_seed(0)
    net_ref = minkunet(name, REF, 3, 20, 3)
    net_gpu = minkunet(name, ME, 3, 20, 3)
    net_gpu.load_state_dict(net_ref.state_dict())
    net_gpu = net_gpu.to(cuda)
    coords = O.surface_cloud(n, seed=3)
    feats = torch.rand(n, 3, generator=torch.Generator().manual_seed(1))
    out_r, loss_r = _run(REF, net_ref, coords, feats, "cpu")
    out_g, loss_g = _run(ME, net_gpu, coords, feats, cuda)
    # stride-1 output rows keep the input order on both sides
    assert torch.equal(out_g.C.cpu(), out_r.C)
    a, b = out_g.F.detach().cpu().numpy(), out_r.F.detach().numpy()
    assert np.abs(a - b).max() / np.abs(b).max() < 1e-3
    assert abs(loss_g - loss_r) / abs(loss_r) < 1e-3
    # gradients of the first and a deep layer (reduction order differs: relative to the max)
    for pname in ("conv0p1s1.kernel", "block4.0.conv1.kernel", "final.kernel"):
        gg = dict(net_gpu.named_parameters())[pname].grad.cpu().numpy()
        gr = dict(net_ref.named_parameters())[pname].grad.numpy()
        # fp32 round-off is amplified through ~30 train-mode batch-norm layers on the way back
        # to the first convolution (observed 4.5e-3 there, <1e-3 at the head); per-op parity
        # is asserted at 1e-5 in test_gpu_parity.py
        tol = 2e-3 if pname == "final.kernel" else 2e-2
        assert np.abs(gg - gr).max() / max(np.abs(gr).max(), 1e-20) < tol, pname
