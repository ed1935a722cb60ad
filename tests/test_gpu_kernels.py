"""Per-kernel parity (GPU): each sm_100a kernel, called through the C ABI, against a plain fp32 torch statement of the
same op on the same (bf16-representable) inputs.  Tolerances are written next to each check."""
import math

import pytest
import torch

from tests.util import assert_close

pytestmark = pytest.mark.gpu

BF16_EPS = 2.0 ** -8           # one bf16 rounding of the output
DEV = "cuda"


def _ops():
    from molly_b200 import ops, _lib
    return ops, _lib


def bf16r(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16)


# ------------------------------------------------------------------------------------------------ GEMM
GEMM_SHAPES = [
    (128, 256, 64),      # one tile, one k-block
    (128, 256, 256),     # k loop
    (256, 512, 128),     # 2x2 tiles
    (300, 320, 320),     # ragged M, N -> partial 128/256 tile, K = 5 k-blocks (ESM-2 t6 shapes)
    (1024, 3840, 1280),  # ESM-2 650M QKV slice
    (77, 96, 64),        # tiny everything (D=96 projector of the tiny fixture)
    (19000, 1280, 512),  # > 148 tiles: persistent loop + TMEM double buffering + phase wrap
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_bias(M, N, K):
    ops, L = _ops()
    a, w = bf16r(M, K, seed=1).to(DEV), bf16r(N, K, scale=0.05, seed=2).to(DEV)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(3)).to(DEV)
    ref = a.float() @ w.float().t() + bias
    out32 = ops.gemm_bf16(a, w, L.EPI_BIAS, bias=bias, out_dtype=torch.float32)
    assert_close(f"gemm_bias_f32 {M}x{N}x{K}", out32, ref, 2e-5)
    out16 = ops.gemm_bf16(a, w, L.EPI_BIAS, bias=bias, out_dtype=torch.bfloat16)
    assert_close(f"gemm_bias_bf16 {M}x{N}x{K}", out16, ref, BF16_EPS)


def test_gemm_nobias_and_qscale():
    ops, L = _ops()
    M, N, K = 256, 384, 128
    a, w = bf16r(M, K, seed=4).to(DEV), bf16r(N, K, scale=0.05, seed=5).to(DEV)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(6)).to(DEV)
    ref = a.float() @ w.float().t()
    assert_close("gemm_nobias", ops.gemm_bf16(a, w, L.EPI_BIAS, out_dtype=torch.float32), ref, 2e-5)
    refq = ref + bias
    refq[:, :128] *= 0.25
    got = ops.gemm_bf16(a, w, L.EPI_BIAS, bias=bias, out_dtype=torch.float32, scale_cols=128, scale=0.25)
    assert_close("gemm_qscale", got, refq, 2e-5)


@pytest.mark.parametrize("heads,d,k,n_seq", [(20, 64, 200, 3), (20, 16, 77, 4), (16, 32, 130, 2)])
def test_gemm_qkv_rope_epilogue(heads, d, k, n_seq):
    """Fused QKV GEMM: bias, q *= d^-1/2 (HF:341), then NeoX rotary on q and k with position = row inside the padded
    sequence (HF:45-54, 81-123); v passes through."""
    ops, L = _ops()
    from oracle.esm_oracle import rotary_tables, rotate_half
    h = heads * d
    M = n_seq * k
    a, w = bf16r(M, h, seed=70).to(DEV), bf16r(3 * h, h, scale=0.05, seed=71).to(DEV)
    bias = torch.randn(3 * h, generator=torch.Generator().manual_seed(72)).to(DEV)
    y = (a.float() @ w.float().t() + bias).cpu().view(n_seq, k, 3, heads, d)
    y[:, :, 0] *= d ** -0.5
    cos, sin = rotary_tables(k, d)
    ref = y.clone()
    for which in (0, 1):
        t = y[:, :, which].permute(0, 2, 1, 3)
        ref[:, :, which] = (t * cos + rotate_half(t) * sin).permute(0, 2, 1, 3)
    cos_t = cos[:, :d // 2].t().contiguous().to(DEV)                  # frequency-major [d/2, k]
    sin_t = sin[:, :d // 2].t().contiguous().to(DEV)
    got = ops.gemm_bf16(a, w, L.EPI_BIAS_ROPE, bias=bias, seq_k=k, scale_cols=h, scale=d ** -0.5,
                        rope_cos_t=cos_t, rope_sin_t=sin_t, rope_cols=2 * h, rope_head_dim=d)
    assert_close(f"gemm_qkv_rope d={d}", got.view(n_seq, k, 3, heads, d), ref, BF16_EPS)


@pytest.mark.parametrize("M,N,K", [(256, 512, 128), (300, 1280, 320), (1024, 5120, 1280)])
def test_gemm_gelu(M, N, K):
    ops, L = _ops()
    a, w = bf16r(M, K, seed=7).to(DEV), bf16r(N, K, scale=0.05, seed=8).to(DEV)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(9)).to(DEV)
    z = a.float() @ w.float().t() + bias
    ref = z * 0.5 * (1.0 + torch.erf(z / math.sqrt(2.0)))          # HF:57-61 exact-erf GELU
    assert_close(f"gemm_gelu {M}x{N}x{K}", ops.gemm_bf16(a, w, L.EPI_BIAS_GELU, bias=bias), ref, BF16_EPS)


@pytest.mark.parametrize("M,N,K", [(256, 256, 512), (300, 320, 1280), (2048, 1280, 5120)])
def test_gemm_residual_inplace(M, N, K):
    ops, L = _ops()
    a, w = bf16r(M, K, seed=10).to(DEV), bf16r(N, K, scale=0.03, seed=11).to(DEV)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(12)).to(DEV)
    x = torch.randn(M, N, generator=torch.Generator().manual_seed(13)).to(DEV)
    ref = a.float() @ w.float().t() + bias + x
    out = ops.gemm_bf16(a, w, L.EPI_BIAS_RESIDUAL, bias=bias, residual=x, out=x)       # in place on the residual stream
    assert out.data_ptr() == x.data_ptr()
    assert_close(f"gemm_residual {M}x{N}x{K}", x, ref, 2e-5)


@pytest.mark.parametrize("M,F,K", [(256, 256, 128), (300, 2048, 512)])
def test_gemm_glu(M, F, K):
    ops, L = _ops()
    a = bf16r(M, K, seed=14).to(DEV)
    w = bf16r(2 * F, K, scale=0.05, seed=15).to(DEV)                      # rows [0,F) = x1, [F,2F) = x2 (NT-v2)
    u = a.float() @ w.float().t()
    ref = torch.nn.functional.silu(u[:, :F]) * u[:, F:]
    w_inter = torch.stack([w[:F], w[F:]], dim=1).reshape(2 * F, K).contiguous()
    assert_close(f"gemm_glu {M}x{F}x{K}", ops.gemm_bf16(a, w_inter, L.EPI_GLU), ref, 1.5 * BF16_EPS)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_gemm_scatter(dtype):
    ops, L = _ops()
    n_seq, k, h, D, B, T, k_cap = 5, 40, 128, 96, 3, 200, 33
    a, w = bf16r(n_seq * k, h, seed=16).to(DEV), bf16r(D, h, scale=0.05, seed=17).to(DEV)
    bias = torch.randn(D, generator=torch.Generator().manual_seed(18)).to(DEV)
    table = torch.tensor([[0, 3], [0, 60], [2, 0], [1, -1], [2, 150]], dtype=torch.int32, device=DEV)
    hs = torch.randn(B, T, D, generator=torch.Generator().manual_seed(19)).to(dtype).to(DEV)
    ref = hs.clone().float()
    y = a.float() @ w.float().t() + bias
    for n, (b, s) in enumerate(table.tolist()):
        if s >= 0:
            ref[b, s + 1:s + 1 + k_cap] = y[n * k:n * k + k_cap]
    untouched = torch.ones(B, T, dtype=torch.bool)
    for n, (b, s) in enumerate(table.tolist()):
        if s >= 0:
            untouched[b, s + 1:s + 1 + k_cap] = False
    before = hs.clone()
    ops.gemm_bf16(a, w, L.EPI_SCATTER, bias=bias, out=hs, seq_table=table, seq_k=k, k_cap=k_cap)
    assert torch.equal(hs.cpu()[untouched], before.cpu()[untouched]), "rows outside the placeholder runs were modified"
    assert_close(f"gemm_scatter {dtype}", hs, ref, BF16_EPS if dtype == torch.bfloat16 else 2e-5)


@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (300, 320, 96), (1000, 1280, 5120), (8192, 3840, 1280), (77, 96, 64)])
def test_linear_wgrad(M, N, K):
    """Autograd of nn.Linear w.r.t. weight and bias: dW = dy^T x (both operands MN-major for the MMA, no transposes)."""
    ops, _ = _ops()
    dy, x = bf16r(M, N, seed=80), bf16r(M, K, seed=81)
    dW, db = ops.linear_wgrad(dy.to(DEV), x.to(DEV))
    assert_close(f"linear_wgrad dW {M}x{N}x{K}", dW, dy.float().t() @ x.float(), 2e-5 * max(1.0, M / 128) ** 0.5)
    assert_close(f"linear_wgrad db {M}x{N}", db, dy.float().sum(0), 1e-5 * max(1.0, M / 128) ** 0.5)


# ------------------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("rows,h,eps", [(1000, 1280, 1e-5), (77, 320, 1e-5), (513, 512, 1e-12), (64, 2560, 1e-12),
                                        (33, 64, 1e-5)])
def test_layernorm(rows, h, eps):
    ops, _ = _ops()
    g = torch.Generator().manual_seed(20)
    x = (torch.randn(rows, h, generator=g) * 3 + 0.5).to(DEV)
    w, b = (1 + 0.1 * torch.randn(h, generator=g)).to(DEV), (0.1 * torch.randn(h, generator=g)).to(DEV)
    ref = torch.nn.functional.layer_norm(x, (h,), w, b, eps)
    assert_close(f"layernorm_f32 h={h}", ops.layernorm(x, w, b, eps, torch.float32), ref, 1e-5)
    assert_close(f"layernorm_bf16 h={h}", ops.layernorm(x, w, b, eps, torch.bfloat16), ref, BF16_EPS)


@pytest.mark.parametrize("rows,h,eps", [(1000, 1280, 1e-5), (77, 320, 1e-5), (513, 1024, 1e-12), (64, 2560, 1e-12),
                                        (33, 64, 1e-5)])
def test_layernorm_backward(rows, h, eps):
    """Autograd of LayerNorm w.r.t. input, weight and bias; the input gradient accumulates into the residual gradient."""
    ops, _ = _ops()
    g = torch.Generator().manual_seed(90)
    x = (torch.randn(rows, h, generator=g) * 3 + 0.5).requires_grad_(True)
    w = (1 + 0.1 * torch.randn(h, generator=g)).requires_grad_(True)
    b = (0.1 * torch.randn(h, generator=g)).requires_grad_(True)
    dy = bf16r(rows, h, seed=91)
    resid = torch.randn(rows, h, generator=g)
    torch.nn.functional.layer_norm(x, (h,), w, b, eps).backward(dy.float())
    d_x = resid.clone().to(DEV)
    d_w, d_b = torch.zeros(h, device=DEV), torch.zeros(h, device=DEV)
    ops.layernorm_bwd(x.detach().to(DEV), dy.to(DEV), w.detach().to(DEV), eps, d_x, True, d_w, d_b)
    assert_close(f"layernorm_bwd dx (+residual) h={h}", d_x, x.grad + resid, 2e-5)
    assert_close(f"layernorm_bwd dgamma h={h}", d_w, w.grad, 1e-4)
    assert_close(f"layernorm_bwd dbeta h={h}", d_b, b.grad, 1e-4)
    d_x2 = torch.empty(rows, h, device=DEV)
    ops.layernorm_bwd(x.detach().to(DEV), dy.to(DEV), w.detach().to(DEV), eps, d_x2, False)
    assert_close(f"layernorm_bwd dx h={h}", d_x2, x.grad, 2e-5)


# ------------------------------------------------------------------------------------------------ rotary
@pytest.mark.parametrize("heads,d,k", [(20, 64, 100), (20, 16, 37), (16, 32, 64), (4, 128, 50)])
def test_rotary(heads, d, k):
    ops, _ = _ops()
    from oracle.esm_oracle import rotary_tables, rotate_half
    n_seq, h = 3, heads * d
    qkv = bf16r(n_seq * k, 3 * h, seed=21).to(DEV)
    cos, sin = rotary_tables(k, d)                                    # [k, d] (cat(freqs, freqs))
    x = qkv.float().cpu().view(n_seq, k, 3, heads, d)
    ref = x.clone()
    for which in (0, 1):
        t = x[:, :, which].permute(0, 2, 1, 3)                        # [n, H, k, d]
        ref[:, :, which] = (t * cos + rotate_half(t) * sin).permute(0, 2, 1, 3)
    got = ops.rotary_(qkv.clone(), k, heads, cos[:, :d // 2].contiguous().to(DEV), sin[:, :d // 2].contiguous().to(DEV))
    assert_close(f"rotary d={d}", got.view(n_seq, k, 3, heads, d), ref, BF16_EPS)
    assert torch.equal(got.view(n_seq, k, 3, heads, d)[:, :, 2].cpu(), qkv.view(n_seq, k, 3, heads, d)[:, :, 2].cpu())


# ------------------------------------------------------------------------------------------------ attention
def _attention_ref(qkv, n_seq, k, heads, key_mask):
    h = qkv.shape[1] // 3
    d = h // heads
    x = qkv.float().view(n_seq, k, 3, heads, d)
    q, kk, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))     # [n, H, k, d]
    s = q @ kk.transpose(2, 3)                                        # scaling 1.0: q is pre-scaled (HF:315,341)
    s = s.masked_fill(~key_mask.bool().view(n_seq, 1, 1, k), float("-inf"))
    return (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(n_seq * k, h)


ATT_CASES = [
    # heads, d, k, valid lengths
    (4, 64, 128, [128, 128]),
    (4, 64, 256, [256, 200, 129, 1]),
    (20, 64, 1024, [1024, 700]),
    (3, 64, 171, [171, 50, 100]),          # k not a multiple of 128 (cfg-4 valid length)
    (20, 16, 512, [512, 33]),              # ESM-2 t6 head_dim 16
    (16, 32, 300, [300, 256, 17]),         # NT-v2-50M head_dim 32
    (2, 128, 256, [256, 171]),             # NT-2.5B head_dim 128
]


@pytest.mark.parametrize("heads,d,k,lens", ATT_CASES)
def test_attention(heads, d, k, lens):
    ops, _ = _ops()
    n_seq, h = len(lens), heads * d
    qkv = bf16r(n_seq * k, 3 * h, seed=22)
    qkv.view(n_seq * k, 3, h)[:, 0] *= d ** -0.5                      # q arrives pre-scaled
    key_mask = torch.zeros(n_seq, k, dtype=torch.uint8)
    for i, ln in enumerate(lens):
        key_mask[i, :ln] = 1
    kv_info = torch.tensor([[ln, ln] for ln in lens], dtype=torch.int32)
    ref = _attention_ref(qkv, n_seq, k, heads, key_mask)
    got = ops.attention(qkv.to(DEV), n_seq, k, heads, kv_info.to(DEV), key_mask.reshape(-1).to(DEV))
    assert_close(f"attention H={heads} d={d} k={k}", got, ref, 1e-2)


@pytest.mark.parametrize("heads,d,k,lens", [(4, 64, 256, [256, 200, 129, 1]), (3, 64, 171, [171, 50, 100]),
                                            (16, 32, 300, [300, 256, 17]), (2, 128, 256, [256, 171])])
def test_attention_log_sum_exp(heads, d, k, lens):
    """The statistic the attention backward starts from: lse2 = log2(sum_j exp(s_j)) per (sequence, head, query row)."""
    import math
    ops, _ = _ops()
    n_seq, h = len(lens), heads * d
    qkv = bf16r(n_seq * k, 3 * h, seed=27)
    qkv.view(n_seq * k, 3, h)[:, 0] *= d ** -0.5
    key_mask = torch.zeros(n_seq, k, dtype=torch.uint8)
    for i, ln in enumerate(lens):
        key_mask[i, :ln] = 1
    kv_info = torch.tensor([[ln, ln] for ln in lens], dtype=torch.int32)
    x = qkv.float().view(n_seq, k, 3, heads, d)
    q, kk = x[:, :, 0].permute(0, 2, 1, 3), x[:, :, 1].permute(0, 2, 1, 3)
    s_ = (q @ kk.transpose(2, 3)).masked_fill(~key_mask.bool().view(n_seq, 1, 1, k), float("-inf"))
    want = torch.logsumexp(s_, -1) / math.log(2.0)                    # [n, H, k]
    out, lse2 = ops.attention_lse(qkv.to(DEV), n_seq, k, heads, kv_info.to(DEV), key_mask.reshape(-1).to(DEV))
    assert_close("attention (lse variant) output", out, _attention_ref(qkv, n_seq, k, heads, key_mask), 1e-2)
    assert float((lse2.cpu() - want).abs().max()) < 2e-3, float((lse2.cpu() - want).abs().max())


@pytest.mark.parametrize("heads,d,k,lens", [(4, 64, 128, [128, 128]), (4, 64, 256, [256, 200, 129, 1]),
                                            (3, 64, 171, [171, 50, 100]), (16, 32, 300, [300, 256, 17]),
                                            (2, 128, 256, [256, 171]), (20, 16, 512, [512, 33]), (20, 64, 1024, [1024, 700])])
def test_attention_backward(heads, d, k, lens):
    """d(q', k', v) of the fused MHA against fp32 autograd of its plain statement (q' = scaled + rotated query as stored)."""
    ops, _ = _ops()
    n_seq, h = len(lens), heads * d
    qkv = bf16r(n_seq * k, 3 * h, seed=41)
    qkv.view(n_seq * k, 3, h)[:, 0] *= d ** -0.5
    key_mask = torch.zeros(n_seq, k, dtype=torch.uint8)
    for i, ln in enumerate(lens):
        key_mask[i, :ln] = 1
    kv_info = torch.tensor([[ln, ln] for ln in lens], dtype=torch.int32)
    d_out = bf16r(n_seq * k, h, seed=42)
    x = qkv.float().requires_grad_(True)
    _attention_ref(x, n_seq, k, heads, key_mask).backward(d_out.float())
    ref = x.grad.view(n_seq * k, 3, h)
    qd, km = qkv.to(DEV), key_mask.reshape(-1).to(DEV)
    out, lse2 = ops.attention_lse(qd, n_seq, k, heads, kv_info.to(DEV), km)
    got = ops.attention_bwd(qd, out, d_out.to(DEV), lse2, n_seq, k, heads, kv_info.to(DEV), km).view(n_seq * k, 3, h)
    for i, name in enumerate(("dq", "dk", "dv")):
        assert_close(f"attention backward {name} H={heads} d={d} k={k}", got[:, i], ref[:, i], 1.5e-2)


def test_attention_backward_interior_pads():
    ops, _ = _ops()
    heads, d, k, n_seq = 4, 64, 300, 3
    h = heads * d
    qkv = bf16r(n_seq * k, 3 * h, seed=43)
    g = torch.Generator().manual_seed(44)
    key_mask = (torch.rand(n_seq, k, generator=g) > 0.3).to(torch.uint8)
    key_mask[0, :] = 1
    key_mask[1, 250:] = 0
    key_mask[2, :130] = 0
    kv_info = torch.tensor([[int(m.nonzero().max()) + 1, int(m.sum())] for m in key_mask], dtype=torch.int32)
    d_out = bf16r(n_seq * k, h, seed=45)
    x = qkv.float().requires_grad_(True)
    _attention_ref(x, n_seq, k, heads, key_mask).backward(d_out.float())
    ref = x.grad.view(n_seq * k, 3, h)
    qd, km = qkv.to(DEV), key_mask.reshape(-1).to(DEV)
    out, lse2 = ops.attention_lse(qd, n_seq, k, heads, kv_info.to(DEV), km)
    got = ops.attention_bwd(qd, out, d_out.to(DEV), lse2, n_seq, k, heads, kv_info.to(DEV), km).view(n_seq * k, 3, h)
    for i, name in enumerate(("dq", "dk", "dv")):
        assert_close(f"attention backward interior pads {name}", got[:, i], ref[:, i], 1.5e-2)


def test_attention_interior_pads():
    """`mask = ids != 1` is not required to be a prefix: interior pad ids must be masked as keys (omics_one.py:70)."""
    ops, _ = _ops()
    heads, d, k, n_seq = 4, 64, 300, 3
    h = heads * d
    qkv = bf16r(n_seq * k, 3 * h, seed=23)
    g = torch.Generator().manual_seed(24)
    key_mask = (torch.rand(n_seq, k, generator=g) > 0.3).to(torch.uint8)
    key_mask[0, :] = 1
    key_mask[1, 250:] = 0
    key_mask[2, :130] = 0                                             # first whole KV block masked
    kv_info = torch.tensor([[int(m.nonzero().max()) + 1, int(m.sum())] for m in key_mask], dtype=torch.int32)
    ref = _attention_ref(qkv, n_seq, k, heads, key_mask)
    got = ops.attention(qkv.to(DEV), n_seq, k, heads, kv_info.to(DEV), key_mask.reshape(-1).to(DEV))
    assert_close("attention interior pads", got, ref, 1e-2)


# ------------------------------------------------------------------------------------------------ embedding gather
@pytest.mark.parametrize("spec_name", ["tiny_esm2", "tiny_ntv1", "tiny_ntv2", "esm2_t6_8m"])
def test_embed(spec_name):
    ops, L = _ops()
    from oracle.esm_oracle import SPECS, esm_embeddings, init_encoder_weights
    from oracle import synth
    from molly_b200.config import EncoderConfig
    from molly_b200.packing import PackedEncoder
    spec = SPECS[spec_name]
    W = init_encoder_weights(spec, 30)
    g = torch.Generator().manual_seed(31)
    k = 70
    rows = []
    for valid in (70, 33, 2, 64):
        rows.append(synth.protein_ids(g, k, valid) if spec.vocab_size == 33 else synth.nucleotide_ids(g, k, valid, spec.vocab_size))
    ids = torch.stack(rows)
    ids[0, 5] = spec.mask_token_id                                   # exercise token-dropout rescale
    ids[0, 9] = spec.mask_token_id
    ids[1, 3] = 1                                                    # interior pad
    ref = esm_embeddings(spec, W, ids, (ids != 1).long())
    c = L.EncoderConfig(hidden_size=spec.hidden_size, vocab_size=spec.vocab_size, pad_token_id=spec.pad_token_id,
                        mask_token_id=spec.mask_token_id,
                        position_type=L.POS_ABSOLUTE if spec.position_embedding_type == "absolute" else L.POS_ROTARY,
                        max_positions=spec.max_position_embeddings, token_dropout=int(spec.token_dropout),
                        emb_layer_norm_before=0)
    we = W["esm.embeddings.word_embeddings.weight"].to(torch.bfloat16).to(DEV)
    pe = W.get("esm.embeddings.position_embeddings.weight")
    pe = None if pe is None else pe.to(torch.bfloat16).to(DEV)
    x, kv_info, key_mask = ops.embed(ids.to(DEV), c, we, pe)
    assert_close(f"embed {spec_name}", x.view(4, k, -1), ref, 1e-6)
    assert key_mask.cpu().view(4, k).equal((ids != 1).to(torch.uint8))
    exp_info = torch.tensor([[int((r != 1).nonzero().max()) + 1, int((r != 1).sum())] for r in ids], dtype=torch.int32)
    assert kv_info.cpu().equal(exp_info)
    assert int(ops.error_flag(torch.device(DEV)).item()) == 0


def test_embed_oov_sets_flag():
    ops, L = _ops()
    ids = torch.tensor([[0, 5, 40, 2, 1, 1]], dtype=torch.int64, device=DEV)       # 40 >= vocab 33
    c = L.EncoderConfig(hidden_size=64, vocab_size=33, pad_token_id=1, mask_token_id=32, position_type=L.POS_ROTARY,
                        max_positions=100, token_dropout=1, emb_layer_norm_before=0)
    we = torch.zeros(33, 64, dtype=torch.bfloat16, device=DEV)
    ops.embed(ids, c, we, None)
    with pytest.raises(AssertionError):
        ops.check_device_errors(torch.device(DEV))
    assert int(ops.error_flag(torch.device(DEV)).item()) == 0


# ------------------------------------------------------------------------------------------------ merge side
def test_placeholder_scan_matches_reference_index_set():
    """Bit-exact: scan(input_ids) == { info.start + 1 + j } (SURVEY.md 0.3 / omics_dataset.py:270-288)."""
    ops, _ = _ops()
    from oracle import synth
    from oracle.esm_oracle import SPECS
    K = 37
    samples = [[("dna", 30), ("protein", 20)], [("protein", 37)], [], [("rna", 5), ("dna", 9), ("protein", 11)],
               [("rna", 37)]]
    for left_pad in (False, True):
        bt = synth.make_batch(5, samples, T=300, D=32, K=K, nt_spec=SPECS["tiny_ntv2"], pr_spec=SPECS["tiny_esm2"],
                              left_pad=left_pad)
        pos, kind, cnt = torch.ops.molly_b200.placeholder_scan(bt.input_ids.to(DEV), *synth.PAD_TOKEN_IDS)
        pos, kind, cnt = pos.cpu(), kind.cpu(), cnt.cpu()
        exp = synth.expected_rows(bt.omic_info_list, K, K, K)
        got = {(b, int(pos[b, c])) for b in range(len(samples)) for c in range(int(cnt[b]))}
        assert got == set(exp.keys())
        for b, infos in enumerate(bt.omic_info_list):
            p = pos[b, :int(cnt[b])].tolist()
            assert p == sorted(p)
            live = [i for i in infos if i["type"] != "pad"]
            assert int(cnt[b]) == K * len(live)
            # run r (K consecutive positions) <-> infos[r] in text order; kinds agree
            for r, info in enumerate(sorted(live, key=lambda i: i["start"])):
                assert p[r * K:(r + 1) * K] == list(range(info["start"] + 1, info["start"] + 1 + K))
                assert set(kind[b, r * K:(r + 1) * K].tolist()) == {synth.KIND_ORDER[info["type"]]}


def test_placeholder_scan_long_rows():
    ops, _ = _ops()
    g = torch.Generator().manual_seed(40)
    B, T = 7, 4099
    ids = torch.randint(0, 1000, (B, T), generator=g)
    m = torch.rand(B, T, generator=g) < 0.37
    kinds = torch.randint(0, 3, (B, T), generator=g)
    pads = torch.tensor([151670, 151673, 151676])
    ids[m] = pads[kinds[m]]
    ids[3] = 5                                                       # a sample with no placeholder at all
    pos, kind, cnt = torch.ops.molly_b200.placeholder_scan(ids.to(DEV), *pads.tolist())
    for b in range(B):
        exp = torch.isin(ids[b], pads).nonzero().flatten()
        assert int(cnt[b]) == len(exp)
        assert pos[b, :len(exp)].cpu().tolist() == exp.tolist()
        assert kind[b, :len(exp)].cpu().tolist() == [pads.tolist().index(int(v)) for v in ids[b][exp]]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_merge_rows(dtype):
    ops, _ = _ops()
    n_seq, k, D, B, T, k_cap = 4, 20, 64, 2, 100, 20
    src = torch.randn(n_seq * k, D, generator=torch.Generator().manual_seed(41)).to(dtype)
    hs = torch.randn(B, T, D, generator=torch.Generator().manual_seed(42)).to(dtype)
    table = torch.tensor([[0, 0], [1, 79], [0, 40], [1, -1]], dtype=torch.int32)
    ref = hs.clone()
    for n, (b, s) in enumerate(table.tolist()):
        if s >= 0:
            ref[b, s + 1:s + 1 + k_cap] = src[n * k:n * k + k_cap]
    got = ops.merge_rows_(hs.to(DEV), src.to(DEV), table.to(DEV), k, k_cap)
    assert torch.equal(got.cpu(), ref)                               # pure copy: bit-exact
