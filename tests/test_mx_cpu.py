"""MXFP8 helpers (ops/mx.py): quantiser oracle, scale-chunk packing round trips."""
import torch

from infomesh_b200.ops import mx as MX


def test_quantize_ref_bounds_and_roundtrip():
    torch.manual_seed(0)
    x = torch.randn(70, 256) * torch.logspace(-3, 3, 70)[:, None]
    q, e = MX.quantize_ref(x)
    assert q.shape == (70, 256) and e.shape == (70, 8) and q.dtype == torch.uint8
    back = MX.dequantize(q, e)
    blk = x.reshape(70, 8, 32)
    amax = blk.abs().amax(-1, keepdim=True)
    # e4m3 keeps 3 mantissa bits: error <= 2^-4 of the block maximum's power-of-two ceiling
    assert ((back.reshape(70, 8, 32) - blk).abs() <= amax * 2 ** -3 + 1e-12).all()
    scale = (e.int() << 23).view(torch.float32)
    assert (amax.squeeze(-1) / scale <= 448.0).all()          # never saturates
    assert (amax.squeeze(-1) / scale > 448.0 / 2 - 1e-3).all()  # and the scale is the smallest such power of two


def test_zero_block_is_finite():
    q, e = MX.quantize_ref(torch.zeros(4, 64))
    assert torch.isfinite(MX.dequantize(q, e)).all() and (q == 0).all()


def test_sfa_pack_roundtrip_and_layout():
    e = torch.arange(300 * 8, dtype=torch.int64).reshape(300, 8).remainder(251).to(torch.uint8)
    ch = MX.pack_sfa(e)
    assert ch.shape == (3, 2, 512)
    assert torch.equal(MX.unpack_sfa(ch, 300), e)
    r, s = 128 + 37 + 64, 5                         # row 229 -> block 1, m1 = 3, m0 = 5;  s = 5 -> kb 1, byte 1
    assert ch[1, 1, 5 * 16 + 3 * 4 + 1] == e[r, s]


def test_sfb_pack_layout():
    e = torch.arange(768 * 24).reshape(768, 24).remainder(241).to(torch.uint8)
    ch = MX.pack_sfb(e)
    assert ch.shape == (6, MX.sfb_chunks(768), 512) and MX.sfb_chunks(768) == 6
    n, s = 5 * 128 + 2 * 32 + 9, 13                 # chunk 5, m1 = 2, m0 = 9; kb 3, byte 1
    assert ch[3, 5, 9 * 16 + 2 * 4 + 1] == e[n, s]
    assert MX.sfb_chunks(384) == 3 and MX.sfb_chunks(2304) == 18 and MX.sfb_chunks(1152) == 9


def test_fused_layernorm_oracle_is_gemm_then_layernorm():
    """``linear_mx_ln_ref`` is what tests/test_gpu_mx.py holds the clustered kernel against: dequantised GEMM + bias + residual,
    then LayerNorm over the full row; the fused kernel only exists for row widths made of 2 or 4 tiles of 192 columns."""
    torch.manual_seed(1)
    m, n, k = 40, 384, 256
    a = MX.quantize_act_ref(torch.randn(m, k))
    w = MX.quantize_weight(torch.randn(n, k) * 0.05)
    bias, res = torch.randn(n), torch.randn(m, n).bfloat16()
    g, b = torch.rand(n) + 0.5, torch.randn(n) * 0.1
    got = MX.linear_mx_ln_ref(a, w, bias, res, g, b, 1e-5)
    y = a.float() @ w.float().T + bias + res.float()
    want = (y - y.mean(-1, keepdim=True)) / torch.sqrt(y.var(-1, unbiased=False, keepdim=True) + 1e-5) * g + b
    assert got.shape == (m, n) and (got - want).abs().max().item() < 1e-4
    assert (MX.linear_mx_ln_ref(a, w, None, None, torch.ones(n), None, 1e-5).mean(-1).abs() < 1e-4).all()   # unit gamma, no beta: zero-mean rows
    assert all(width % 192 == 0 and width // 192 in (2, 4) for width in MX.FUSED_LN_WIDTHS)
