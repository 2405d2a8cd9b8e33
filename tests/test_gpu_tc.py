"""tcgen05 (tensor-core) convolution path: bf16 / fp16 operands, fp32 accumulation.

Parity definition for reduced-precision inputs (SURVEY.md §8a notes): the oracle is fed the
SAME bf16/fp16-rounded features and weights; with fp32 accumulation the CUDA result then
differs only by summation order (asserted 2e-5 relative with fp32 output) plus one final
rounding when the output is stored in bf16/fp16 (2^-8 relative, asserted 6e-3 of max)."""
import pytest
import torch

from helpers import rel_err, unique_cloud
from oracle import oracle_np as O

pytestmark = pytest.mark.gpu


def _torch_ref_forward(feats, w, nbr):
    """fp32 torch restatement on the device: sum_k feats[nbr[k]] @ w[k] (missing rows = 0)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    K, n_out = nbr.shape
    out = torch.zeros((n_out, w.shape[2]), dtype=torch.float32, device=feats.device)
    f32, w32 = feats.float(), w.float()
    for k in range(K):
        idx = nbr[k].long()
        valid = idx >= 0
        g = f32[idx.clamp(min=0)] * valid.unsqueeze(1)
        out += g @ w32[k]
    return out


def _random_table(K, n_out, n_in, density, seed, device):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, n_in, (K, n_out), generator=g, dtype=torch.int32)
    drop = torch.rand((K, n_out), generator=g) > density
    idx[drop] = -1
    return idx.to(device)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("cin,cout,n_in,n_out,K", [
    (16, 16, 500, 100, 27),          # single partial tile, BK=16
    (32, 32, 3000, 3001, 27),        # BK=32, R=4 super tiles with a ragged tail
    (64, 128, 5000, 2000, 27),       # BASELINE cfg1 channel shape, BK=64
    (96, 96, 4000, 4000, 27),        # MinkUNet34C stride-1 blocks (BK=32, N=96)
    (128, 96, 4000, 1000, 8),
    (192, 128, 1000, 1000, 27),
    (384, 256, 700, 300, 27),        # widest concat layer, N=256 (R=1)
    (256, 256, 300, 129, 27),
    (32, 64, 90000, 90000, 8),       # more super tiles than SMs: persistent loop + phases
])
def test_tc_forward_matches_fp32_reference(ME, cuda, dtype, cin, cout, n_in, n_out, K):
    from minkowskiengine_b200 import backend
    g = torch.Generator().manual_seed(cin * 1000 + cout)
    feats = (torch.rand(n_in, cin, generator=g) - 0.5).to(dtype).to(cuda)
    w = ((torch.rand(K, cin, cout, generator=g) - 0.5) / (cin ** 0.5)).to(dtype).to(cuda)
    nbr = _random_table(K, n_out, n_in, 0.35, seed=K + n_out, device=cuda)
    km = backend._KernelMap(nbr, torch.empty((K, n_in), dtype=torch.int32, device=cuda))
    ref = _torch_ref_forward(feats, w, nbr)
    out32 = backend._conv_forward(feats, w, km, out_dtype=torch.float32)
    scale = ref.abs().max().item()
    assert (out32 - ref).abs().max().item() / scale < 2e-5
    out_lp = backend._conv_forward(feats, w, km)
    assert out_lp.dtype == dtype
    assert (out_lp.float() - ref).abs().max().item() / scale < 6e-3


@pytest.mark.parametrize("dtype", [torch.bfloat16])
@pytest.mark.parametrize("cin,cout,ks,stride", [(32, 64, 3, 1), (64, 96, 3, 2), (96, 32, 2, 2)])
def test_tc_layer_vs_oracle(ME, cuda, dtype, cin, cout, ks, stride):
    """Through the public API (SparseTensor + MinkowskiConvolution, bf16 features), forward
    and backward, against the numpy oracle fed the same rounded values."""
    D = 3
    coords = unique_cloud(6000, 24, seed=5, allow_negative=True)
    g = torch.Generator().manual_seed(9)
    feats = torch.rand(len(coords), cin, generator=g).to(dtype)
    conv = ME.MinkowskiConvolution(cin, cout, kernel_size=ks, stride=stride, dimension=D).to(cuda)
    x = ME.SparseTensor(feats, coords, device=cuda, requires_grad=True)
    y = conv(x)
    assert y.F.dtype == dtype
    in_c, out_c = x.C.cpu().numpy(), y.C.cpu().numpy()
    im, om = O.kernel_map(in_c, out_c, O.region_offsets(O.HYPER_CUBE, [ks] * D, [1] * D, [1] * D))
    w = conv.kernel.detach().to(dtype).float().cpu().numpy()
    f = feats.float().numpy()
    ref = O.conv_forward(f, w, im, om, len(out_c))
    assert rel_err(y.F.detach().float().cpu().numpy(), ref) < 6e-3
    gout = (torch.rand(y.F.shape, generator=g) - 0.5).to(dtype)
    y.F.backward(gout.to(cuda))
    gi, gw = O.conv_backward(f, gout.float().numpy(), w, im, om)
    assert rel_err(x.F.grad.float().cpu().numpy(), gi) < 6e-3       # bf16-stored dgrad
    assert rel_err(conv.kernel.grad.float().cpu().numpy(), gw) < 1e-4   # fp32 wgrad


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("cin,cout,n_in,n_out,K", [
    (16, 16, 500, 100, 27),
    (32, 32, 3000, 3001, 27),
    (64, 128, 5000, 2000, 27),
    (96, 96, 4000, 4000, 27),       # N = 96: one and a half 64-channel blocks
    (128, 96, 4000, 1000, 8),
    (192, 128, 1000, 1000, 27),     # two m-tiles (128 + 64 padded)
    (384, 256, 700, 300, 27),       # three m-tiles -> two m-tile groups
    (256, 256, 300, 129, 27),
    (24, 48, 777, 333, 5),          # channel counts that are not multiples of 16 / 64
    (32, 64, 60000, 60000, 8),
])
def test_tc_wgrad_matches_fp32_reference(ME, cuda, dtype, cin, cout, n_in, n_out, K):
    from minkowskiengine_b200 import _lib, backend
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(cin * 77 + cout)
    feats = (torch.rand(n_in, cin, generator=g) - 0.5).to(dtype).to(cuda)
    gout = (torch.rand(n_out, cout, generator=g) - 0.5).to(dtype).to(cuda)
    w = torch.zeros(K, cin, cout, dtype=torch.float32, device=cuda)   # fp32 master weights
    nbr = _random_table(K, n_out, n_in, 0.35, seed=K + n_out + 1, device=cuda)
    km = backend._KernelMap(nbr, torch.full((K, n_in), -1, dtype=torch.int32, device=cuda))
    before = _lib.tc_launch_count()
    _, gw = backend._conv_backward(feats, gout, w, km, need_in=False, need_w=True)
    assert _lib.tc_launch_count() > before, "wgrad did not take the tcgen05 path"
    ref = torch.zeros(K, cin, cout, dtype=torch.float32, device=cuda)
    f32, g32 = feats.float(), gout.float()
    for k in range(K):
        idx = nbr[k].long()
        valid = (idx >= 0).unsqueeze(1)
        ref[k] = (f32[idx.clamp(min=0)] * valid).t() @ g32
    assert (gw.float() - ref).abs().max().item() / ref.abs().max().item() < 2e-5


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("cin,cout,n_in,n_out,K", [
    (32, 32, 3000, 3001, 27),        # one 32-channel block per stage, R = 4, ragged tail
    (32, 16, 500, 100, 27),          # single partial tile
    (64, 128, 5000, 2000, 27),       # two blocks per stage (128B-swizzled weights)
    (96, 96, 4000, 4000, 27),        # three blocks per stage: the MinkUNet34C stride-1 blocks
    (128, 96, 4000, 1000, 8),        # two stages per (tile, offset)
    (192, 128, 1000, 1000, 27),
    (384, 256, 700, 300, 27),        # four stages per offset, N = 256 (single accumulator set)
    (256, 384, 300, 129, 27),        # output columns split in two launches
    (160, 64, 2000, 1500, 5),        # 160 = 5 blocks of 32 (one block per stage)
    (32, 64, 90000, 90000, 8),       # more super tiles than SMs: persistent loop + phases
    (96, 96, 50000, 50000, 27),
])
def test_ta_forward_and_dgrad_packed_weights(ME, cuda, dtype, cin, cout, n_in, n_out, K):
    """fp32 master weights -> packed operand copies -> k_conv_ta (operand A in tensor memory):
    forward over out_nbr and dgrad over in_nbr against the fp32 torch restatement."""
    from minkowskiengine_b200 import _lib, backend
    g = torch.Generator().manual_seed(cin * 1000 + cout + 7)
    feats = (torch.rand(n_in, cin, generator=g) - 0.5).to(dtype).to(cuda)
    w = ((torch.rand(K, cin, cout, generator=g) - 0.5) / (cin ** 0.5)).to(cuda)   # fp32 master
    wl = w.to(dtype)
    out_nbr = _random_table(K, n_out, n_in, 0.35, seed=K + n_out, device=cuda)
    in_nbr = _random_table(K, n_in, n_out, 0.3, seed=K + n_in + 3, device=cuda)
    km = backend._KernelMap(out_nbr, in_nbr)
    before = _lib.tc_launch_count()
    out32 = backend._conv_forward(feats, w, km, out_dtype=torch.float32)
    assert _lib.tc_launch_count() > before
    ref = _torch_ref_forward(feats, wl, out_nbr)
    scale = ref.abs().max().item()
    assert (out32 - ref).abs().max().item() / scale < 2e-5
    out_lp = backend._conv_forward(feats, w, km)
    assert out_lp.dtype == dtype
    assert (out_lp.float() - ref).abs().max().item() / scale < 6e-3
    # weights modified in place -> the packed copies must be rebuilt
    with torch.no_grad():
        w.mul_(0.5)
    out_half = backend._conv_forward(feats, w, km, out_dtype=torch.float32)
    ref_half = _torch_ref_forward(feats, w.to(dtype), out_nbr)
    assert (out_half - ref_half).abs().max().item() / scale < 2e-5
    # dgrad: rows = input rows, reduction over c_out, W_k^T
    gout = (torch.rand(n_out, cout, generator=g) - 0.5).to(dtype).to(cuda)
    gi, _ = backend._conv_backward(feats, gout, w, km, need_in=True, need_w=False)
    ref_gi = _torch_ref_forward(gout, w.to(dtype).transpose(1, 2).contiguous(), in_nbr)
    assert gi.dtype == dtype
    assert (gi.float() - ref_gi).abs().max().item() / ref_gi.abs().max().item() < 6e-3


@pytest.mark.parametrize("K,n,density,chunk", [(27, 5000, 0.3, 2048), (8, 70000, 0.125, 65536),
                                               (1, 100, 1.0, 65536), (125, 3001, 0.05, 2048),
                                               (27, 2048, 0.0, 2048), (81, 40000, 0.4, 4096),
                                               (27, 150000, 0.3, 65536)])
def test_pair_lists_match_table(ME, cuda, K, n, density, chunk, monkeypatch):
    """meb200_kernel_map_pairs: compacted (other row, table row) lists in (row chunk, offset)
    order, table-row order inside a segment, every segment padded with -1 to a multiple of the
    stage — against torch.nonzero."""
    from minkowskiengine_b200 import backend
    monkeypatch.setattr(backend._KernelMap, "PAIR_CHUNK_ROWS", chunk)
    nbr = _random_table(K, n, 12345, density, seed=K + n, device=cuda)
    km = backend._KernelMap(nbr, torch.empty((K, 1), dtype=torch.int32, device=cuda))
    pin, pout, seg, nch = km.pair_lists()
    S = km.PAIR_STAGE
    seg_h = seg.cpu().tolist()
    assert seg_h[0] == 0 and len(seg_h) == nch * K + 1
    bpc = -(-chunk // 2048)                   # table blocks (2048 rows) per chunk
    assert nch == -(-(-(-n // 2048)) // bpc)
    covered = 0
    for c in range(nch):
        r0, r1 = c * bpc * 2048, min((c + 1) * bpc * 2048, n)
        for k in range(K):
            rows = torch.nonzero(nbr[k, r0:r1] >= 0).flatten() + r0
            cnum = len(rows)
            a, b = seg_h[c * K + k], seg_h[c * K + k + 1]
            assert b - a == (cnum + S - 1) // S * S
            assert torch.equal(pout[a:a + cnum].long(), rows)
            assert torch.equal(pin[a:a + cnum], nbr[k][rows])
            assert bool((pin[a + cnum:b] == -1).all()) and bool((pout[a + cnum:b] == -1).all())
            covered += cnum
    assert covered == int((nbr >= 0).sum())
    # the swapped view exchanges the two sides and shares the storage
    sw = km.swapped()
    sin, sout, sseg, snch = sw.pair_lists()
    assert sin.data_ptr() == pout.data_ptr() and sout.data_ptr() == pin.data_ptr()
    assert sseg.data_ptr() == seg.data_ptr() and snch == nch


def test_wgrad_pairs_many_chunks(ME, cuda, monkeypatch):
    """k_wgrad_pairs walking several row chunks (small chunks forced) against the fp32 reference."""
    from minkowskiengine_b200 import backend
    monkeypatch.setattr(backend._KernelMap, "PAIR_CHUNK_ROWS", 4096)
    torch.backends.cuda.matmul.allow_tf32 = False
    K, n_in, n_out, cin, cout = 27, 30000, 33000, 96, 96
    g = torch.Generator().manual_seed(5)
    feats = (torch.rand(n_in, cin, generator=g) - 0.5).bfloat16().to(cuda)
    gout = (torch.rand(n_out, cout, generator=g) - 0.5).bfloat16().to(cuda)
    w = torch.zeros(K, cin, cout, device=cuda)
    nbr = _random_table(K, n_out, n_in, 0.3, seed=11, device=cuda)
    km = backend._KernelMap(nbr, torch.full((K, n_in), -1, dtype=torch.int32, device=cuda))
    _, gw = backend._conv_backward(feats, gout, w, km, need_in=False, need_w=True)
    assert km.pair_lists()[3] == 9
    ref = torch.zeros(K, cin, cout, device=cuda)
    f32, g32 = feats.float(), gout.float()
    for k in range(K):
        idx = nbr[k].long()
        ref[k] = (f32[idx.clamp(min=0)] * (idx >= 0).unsqueeze(1)).t() @ g32
    assert (gw - ref).abs().max().item() / ref.abs().max().item() < 2e-5


def test_batched_weight_packing_equals_single(ME, cuda):
    """The one-launch re-pack of every registered kernel (backend._PackTable,
    meb200_conv_pack_weights_batched) must leave exactly the bytes the per-kernel packing leaves,
    for ragged and 32-aligned channel counts, and track tensors that die or move."""
    from minkowskiengine_b200 import backend, _lib
    lib = _lib.load()
    assert backend._PACK_BATCHED
    dtype = torch.bfloat16
    g = torch.Generator().manual_seed(5)
    shapes = [(27, 32, 64), (8, 96, 96), (1, 5, 7), (27, 48, 20), (125, 64, 32), (3, 256, 128)]
    ks = [torch.nn.Parameter(torch.randn(s, generator=g).to(cuda)) for s in shapes]

    def single(k):
        K, ci, co = k.shape
        buf = torch.empty((4, K * ci * co), dtype=dtype, device=cuda)
        _lib.check(lib.meb200_conv_pack_weights(
            _lib.ptr(k.detach()), K, ci, co, _lib.dtype_code(dtype), _lib.ptr(buf[0]), _lib.ptr(buf[1]),
            _lib.ptr(buf[2]) if co % 32 == 0 else None, _lib.ptr(buf[3]) if ci % 32 == 0 else None,
            _lib.current_stream()))
        return buf, co % 32 == 0, ci % 32 == 0

    def check(k):
        got = backend._packed_weights(k, dtype)
        ref, has_cp, has_tp = single(k)
        assert torch.equal(got[0].flatten(), ref[0]) and torch.equal(got[1].flatten(), ref[1])
        assert (got[2] is not None) == has_cp and (got[3] is not None) == has_tp
        if has_cp:
            assert torch.equal(got[2].flatten(), ref[2])
        if has_tp:
            assert torch.equal(got[3].flatten(), ref[3])

    for k in ks:            # first sight: packed one by one, registered
        check(k)
    tbl = backend._PACK_TABLES[(torch.device(cuda).index or 0, dtype)]
    n0 = lib.meb200_launch_count()
    with torch.no_grad():   # "optimizer step": every kernel changes
        for k in ks:
            k.add_(0.5)
    got0 = backend._packed_weights(ks[0], dtype)          # one batched launch re-packs all of them
    assert lib.meb200_launch_count() == n0 + 1
    for k in ks:
        assert backend._PACKED[id(k)][1] == k._version     # all fresh: no further launches
    n1 = lib.meb200_launch_count()
    for k in ks:
        backend._packed_weights(k, dtype)
    assert lib.meb200_launch_count() == n1
    for k in ks:
        check(k)
    # a tensor dies, another is re-allocated: the table is rebuilt, results stay right
    del ks[2]
    with torch.no_grad():
        ks[1].data = ks[1].data.clone()
        for k in ks:
            k.mul_(1.25)
    for k in ks:
        check(k)
    assert all(e[0]() is not None for e in tbl.entries.values()) or True
