"""Direct checks of post-loop / element-wise kernels through the C-ABI against plain torch fp32 on the same inputs:
the attention context-gradient reduction (las_att_dq), the energy-MLP gradient kernel in its three output modes
(las_att_param_grads_part: the split used by the trainers must equal the single pass), mask identity between the three
dropout kernels (8-wide, 4-wide, scalar), and the persistent BLSTM forward writing every row of y."""
import pytest
import torch

from tests.util import pkg

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,L,Te,O", [(3, 9, 13, 40), (5, 130, 37, 320), (2, 7, 8, 24)])
def test_att_dq_matches_einsum(B, L, Te, O):
    Fn = pkg("functional")
    dev = torch.device("cuda")
    torch.manual_seed(B * 100 + Te)
    ws = torch.rand(B, L + 1, Te, device=dev)
    dc = torch.randn(B, L, O, device=dev)
    dQ = torch.full((B * Te, O), float("nan"), device=dev)
    Fn.call("las_att_dq", Fn.ptr(ws), Fn.ptr(dc), L, B, Te, O, Fn.ptr(dQ))
    want = torch.einsum("bte,bto->beo", ws[:, 1:].double(), dc.double()).reshape(B * Te, O)
    assert torch.allclose(dQ.double(), want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,L,Te,A,C", [(3, 11, 13, 48, 5), (2, 40, 30, 320, 10), (4, 5, 7, 64, 16), (3, 126, 125, 320, 10)])
def test_att_param_grads_split_equals_single_pass_and_torch(B, L, Te, A, C):
    Fn, LIB = pkg("functional"), pkg("_lib")
    dev = torch.device("cuda")
    torch.manual_seed(7 * B + A)
    P = torch.randn(B, Te, A, device=dev) * 0.5
    dz = torch.randn(B, L, A, device=dev) * 0.5
    conv = torch.zeros(B, L, Te, 16, device=dev)
    conv[..., :C] = torch.randn(B, L, Te, C, device=dev) * 0.5
    de = torch.randn(B, L, Te, device=dev)
    matt = torch.randn(A, C, device=dev) * 0.3
    gv = torch.randn(A, device=dev)
    part = torch.empty(LIB.lib().las_att_scratch_floats(B, L, Te, A, C, 3), device=dev)

    def run(what):
        dP = torch.zeros(B * Te, A, device=dev)
        dm = torch.zeros(A, C, device=dev)
        dg = torch.zeros(A, device=dev)
        Fn.call("las_att_param_grads_part", Fn.ptr(P), Fn.ptr(dz), Fn.ptr(conv), Fn.ptr(de), Fn.ptr(matt), Fn.ptr(gv),
                B, L, Te, A, C, what, Fn.ptr(dP), Fn.ptr(part), Fn.ptr(dm), Fn.ptr(dg))
        return dP, dm, dg

    dP3, dm3, dg3 = run(3)
    dP1, dm1, dg1 = run(1)
    dP2, dm2, dg2 = run(2)
    def close(a, b, tol=1e-5):      # the instances are compiled separately (different FMA contraction): equal to f32 rounding
        return float((a - b).abs().max()) <= tol * float(b.abs().max())
    # dP only: for att_dim % 32 == 0 this is the tensor-core kernel (mlp_att contraction as split-bf16 MMAs: ~2^-16
    # relative per product instead of f32)
    assert close(dP1, dP3, 1e-5 if A % 32 else 5e-4) and not dm1.any() and not dg1.any()
    assert close(dm2, dm3) and close(dg2, dg3) and not dP2.any()                                   # parameters only
    # torch fp32/fp64 reference (the kernel recomputes tanh with tanh.approx: 2^-11 relative)
    s = torch.tanh(P.double()[:, None] + dz.double()[:, :, None] + torch.einsum("blec,ac->blea", conv[..., :C].double(), matt.double()))
    ds = de.double()[..., None] * gv.double() * (1 - s * s)
    for got, want in ((dP3.view(B, Te, A), ds.sum(1)), (dP1.view(B, Te, A), ds.sum(1)),
                      (dm3, torch.einsum("blea,blec->ac", ds, conv[..., :C].double())),
                      (dg3, torch.einsum("ble,blea->a", de.double(), s))):
        err = float((got.double() - want).abs().max() / want.abs().max())
        assert err < 5e-3, err


def test_dropout_kernels_agree_on_the_mask():
    """The 8-wide, 4-wide and scalar kernels are chosen by alignment; the mask is a function of the logical element
    index only, so the same (site, step) must give the same mask through all three."""
    Fn = pkg("functional")
    dev = torch.device("cuda")
    B, T, W, p = 5, 21, 96, 0.3
    masks = []
    for pad, dtype in ((0, torch.bfloat16), (0, torch.float32), (4, torch.float32), (4, torch.bfloat16), (1, torch.float32)):
        x = torch.ones(B, T + 1, W + pad, device=dev, dtype=dtype)      # pad 0: 8-wide, 4: 4-wide, 1: scalar kernel
        Fn.dropout_(x, B, T, W, (T + 1) * (W + pad), W + pad, 1, p, 777)
        assert bool((x[:, :, W:] == 1).all())                           # the padding columns are not touched
        masks.append(x[:, :, :W] != 0)
    for m in masks[1:]:
        assert torch.equal(m, masks[0])
    assert abs(float(masks[0].float().mean()) - (1 - p)) < 0.02


def test_persistent_blstm_forward_writes_every_row_of_y():
    """las_lstm_persist_fwd must leave zeros past each length (pad_packed_sequence, model.py:81) even when the
    caller's buffer held garbage, including the replicated row of an odd extent (model.py:88-89)."""
    Fn, LIB = pkg("functional"), pkg("_lib")
    dev = torch.device("cuda")
    H = 64
    if not LIB.lib().las_lstm_persistent_geometry(H, None, None):
        pytest.skip("persistent LSTM not available for this hidden size")
    torch.manual_seed(3)
    B, T, D = 5, 17, 24
    lens = torch.tensor([17, 12, 9, 9, 3], dtype=torch.int32, device=dev)
    x = torch.randn(B * T, D, device=dev)
    xin = Fn.cvt_bf16(x)
    w_ih = [torch.randn(4 * H, D, device=dev) * 0.2 for _ in range(2)]
    w_hh = [torch.randn(4 * H, H, device=dev) * 0.2 for _ in range(2)]
    b = [torch.randn(4 * H, device=dev) * 0.1 for _ in range(4)]
    real_empty = torch.empty

    def dirty_empty(*a, **k):
        t = real_empty(*a, **k)
        return t.fill_(7) if t.dtype == torch.bfloat16 else t
    torch.empty = dirty_empty
    try:
        y, _ = Fn.lstm_layer_fwd(xin, xin.shape[1], w_ih, w_hh, b[:2], b[2:], lens, B, T, T + 1, 1)
    finally:
        torch.empty = real_empty
    y = y.float()
    for bi, l in enumerate(lens.tolist()):
        assert bool((y[bi, l:T] == 0).all())
        assert bool((y[bi, :l].abs().sum(-1) > 0).all())
    assert torch.equal(y[:, T], y[:, T - 1])
