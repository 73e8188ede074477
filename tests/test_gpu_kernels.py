"""GPU: building-block kernels through the C ABI vs plain PyTorch fp32 references of the same op."""
import pytest
import torch

from embodied_captioning_b200 import ops

pytestmark = pytest.mark.gpu


def _rand_bf16(shape, dev, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).to(dev)


GEMM_SHAPES = [
    (128, 256, 64), (128, 256, 128), (256, 512, 1024), (257, 384, 128), (1000, 3072, 1024), (4, 768, 768),
    (50, 128, 3072), (1024, 768, 3072), (300, 1536, 768), (130, 64, 64), (513, 1000, 592), (2570, 1024, 4096),
    (3855, 3072, 1024),
    # one wave of 128 x 192 tiles instead of a ragged second wave of 128 x 128 (bf16 / GELU epilogues); ragged M
    (1001, 3072, 768), (700, 4608, 512),
    # enough 256 x 256 tiles for the CTA-pair (cta_group::2) kernel; ragged M, N and K tails
    (20000, 1024, 1024), (19001, 1000, 592), (9733, 2304, 768),
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_bf16_epilogues(device, M, N, K):
    a = _rand_bf16((M, K), device, seed=1)
    w = _rand_bf16((N, K), device, K ** -0.5, seed=2)
    bias = torch.randn(N, device=device)
    res = torch.randn(M, N, device=device)
    ref = a.float() @ w.float().t() + bias
    out = ops.gemm(a, w, bias, ops.EPI_BF16)
    ops.check_watchdog(device)
    err = (out.float() - ref).abs().max().item()
    assert err < 0.05, f"bf16 epilogue max err {err}"
    assert torch.equal(out, ref.to(torch.bfloat16)) or (out.float() - ref).abs().max().item() < 0.05
    out2 = ops.gemm(a, w, bias, ops.EPI_F32)
    assert (out2 - ref).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())
    out3 = ops.gemm(a, w, bias, ops.EPI_F32_RESID, residual=res)
    assert (out3 - (ref + res)).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())
    out4 = ops.gemm(a, w, bias, ops.EPI_BF16_GELU)
    assert (out4.float() - torch.nn.functional.gelu(ref)).abs().max().item() < 0.05
    out5 = ops.gemm(a, w, None, ops.EPI_F32)
    assert (out5 - (ref - bias)).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())
    ops.check_watchdog(device)


def test_gemm_matches_simt_crosscheck(device):
    a = _rand_bf16((300, 592), device, seed=3)
    w = _rand_bf16((1024, 592), device, 0.05, seed=4)
    t = ops.gemm(a[:, :588].contiguous() if False else a, w, None, ops.EPI_F32)
    s = ops.gemm(a, w, None, ops.EPI_F32, simt_reference=True)
    assert (t - s).abs().max().item() < 1e-3


def test_gemm_k_tail_reads_zero_beyond_k(device):
    """K = 588 inside rows of 592 elements (the ViT-L/14 patch GEMM): the TMA box over-reads columns 588..639,
    which must come back as zeros although the padding columns hold garbage."""
    a = _rand_bf16((257, 592), device, seed=5)
    w = _rand_bf16((256, 592), device, 0.05, seed=6)
    a[:, 588:] = 1000.0
    w[:, 588:] = 1000.0
    from embodied_captioning_b200 import _native as nat
    out = torch.empty((257, 256), device=device, dtype=torch.float32)
    nat.check(nat.lib().ec_gemm_bf16(a.data_ptr(), 592, w.data_ptr(), 592, 257, 256, 588, None, None, 0, out.data_ptr(),
                                     256, ops.EPI_F32, 0, torch.cuda.current_stream().cuda_stream))
    ref = a[:, :588].float() @ w[:, :588].float().t()
    assert (out - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("M,N,K", [(64, 1024, 256), (1024, 49408, 768), (7, 2048, 512)])
def test_gemm_argmax(device, M, N, K):
    a = _rand_bf16((M, K), device, seed=7)
    w = _rand_bf16((N, K), device, K ** -0.5, seed=8)
    arg, logits = ops.gemm_argmax(a, w)
    ops.check_watchdog(device)
    ref = a.float() @ w.float().t()
    assert (logits - ref).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())
    assert torch.equal(arg.long(), logits.argmax(dim=-1))
    top2 = ref.topk(2, dim=-1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-3
    assert torch.equal(arg.long()[clear], ref.argmax(dim=-1)[clear])


@pytest.mark.parametrize("rows,width", [(1, 128), (257, 1024), (1000, 768), (33, 512), (5, 256)])
def test_layernorm(device, rows, width):
    x = torch.randn(rows, width, device=device) * 3 + 1
    g = torch.randn(width, device=device)
    b = torch.randn(width, device=device)
    ref = torch.nn.functional.layer_norm(x, (width,), g, b, 1e-5)
    y32 = ops.layernorm(x, g, b, out_bf16=False)
    assert (y32 - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())
    y16 = ops.layernorm(x, g, b, out_bf16=True)
    assert torch.equal(y16, y32.to(torch.bfloat16))


@pytest.mark.parametrize("B,H,Sq,Skv,D", [(3, 16, 257, 257, 64), (2, 12, 50, 50, 64), (2, 8, 256, 257, 96),
                                          (2, 8, 256, 50, 64), (1, 8, 256, 50, 32), (5, 2, 50, 50, 64),
                                          (1, 1, 1, 1, 64), (2, 4, 17, 33, 64)])
def test_attention_full(device, B, H, Sq, Skv, D):
    C = H * D
    q = _rand_bf16((B, Sq, C), device, seed=11)
    k = _rand_bf16((B, Skv, C), device, seed=12)
    v = _rand_bf16((B, Skv, C), device, seed=13)
    out = ops.attention(q, k, v, H)
    qf = q.float().view(B, Sq, H, D).transpose(1, 2)
    kf = k.float().view(B, Skv, H, D).transpose(1, 2)
    vf = v.float().view(B, Skv, H, D).transpose(1, 2)
    p = torch.softmax(qf @ kf.transpose(-1, -2) / D ** 0.5, dim=-1)
    ref = (p @ vf).transpose(1, 2).reshape(B, Sq, C)
    err = (out.float() - ref).abs().max().item()
    assert err < 0.03, f"attention max err {err}"
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    assert rel < 1e-2, f"attention rel err {rel}"


def test_attention_strided_fused_qkv_and_shared_queries(device):
    """The two layouts the image tower uses: q/k/v as column slices of one [B,S,3C] buffer, and batch-shared
    pooler queries."""
    B, S, H, D = 2, 50, 4, 64
    C = H * D
    qkv = _rand_bf16((B, S, 3 * C), device, seed=21)
    out = ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], H)
    ref = ops.attention(qkv[:, :, :C].contiguous(), qkv[:, :, C:2 * C].contiguous(), qkv[:, :, 2 * C:].contiguous(), H)
    assert torch.equal(out, ref)
    q = _rand_bf16((16, C), device, seed=22)
    out2 = ops.attention(q, qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], H)
    ref2 = ops.attention(q.unsqueeze(0).expand(B, -1, -1).contiguous(), qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], H)
    assert torch.equal(out2, ref2)


@pytest.mark.parametrize("B,H,S", [(3, 16, 257), (2, 12, 50), (2, 4, 256), (2, 4, 258), (5, 3, 130), (1, 1, 17),
                                   (1, 2, 1), (300, 2, 257), (1, 1, 257), (149, 1, 257), (37, 16, 257)])
def test_attention_tcgen05_fused_qkv(device, B, H, S):
    """ViT-block attention on tcgen05/TMEM (S = QK^T, P and O in tensor memory) vs an fp32 reference and vs the
    mma.sync kernel, on the fused [B, S, 3*H*64] layout of the image tower."""
    D = 64
    C = H * D
    qkv = _rand_bf16((B, S, 3 * C), device, seed=31)
    out = ops.attention_fused_qkv(qkv, H, impl=1)
    ops.check_watchdog(device)
    whole_row = ops.attention_fused_qkv(qkv, H, impl=2)     # 257 tokens: impl 1 is the block-wise kernel, 2 the whole-row one
    ops.check_watchdog(device)
    assert (out.float() - whole_row.float()).abs().max().item() < 0.03
    legacy = ops.attention_fused_qkv(qkv, H, impl=0)
    qf = qkv[:, :, :C].float().view(B, S, H, D).transpose(1, 2)
    kf = qkv[:, :, C:2 * C].float().view(B, S, H, D).transpose(1, 2)
    vf = qkv[:, :, 2 * C:].float().view(B, S, H, D).transpose(1, 2)
    p = torch.softmax(qf @ kf.transpose(-1, -2) / D ** 0.5, dim=-1)
    ref = (p @ vf).transpose(1, 2).reshape(B, S, C)
    err = (out.float() - ref).abs().max().item()
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    rel_legacy = ((legacy.float() - ref).norm() / ref.norm()).item()
    print(f"tcgen05 attention B={B} H={H} S={S}: max err {err:.4f}, rel {rel:.2e} (mma.sync kernel rel {rel_legacy:.2e})")
    assert err < 0.03 and rel < 1e-2


@pytest.mark.parametrize("first_big", [100, 200, 256])
def test_attention_tcgen05_257_rescale_path(device, first_big):
    """The block-wise 257-token kernel keeps block 0's row max as the softmax reference unless a later key block exceeds
    it by more than 2^8; keys >= `first_big` are scaled up so that (almost) every row takes the rescale path in block 1
    and / or block 2 (key 256 alone for first_big = 256)."""
    B, H, S, D = 5, 3, 257, 64
    C = H * D
    qkv = _rand_bf16((B, S, 3 * C), device, seed=77).float()
    qkv[:, first_big:, C:2 * C] *= 9.0
    qkv = qkv.to(torch.bfloat16)
    out = ops.attention_fused_qkv(qkv, H, impl=1)
    ops.check_watchdog(device)
    qf = qkv[:, :, :C].float().view(B, S, H, D).transpose(1, 2)
    kf = qkv[:, :, C:2 * C].float().view(B, S, H, D).transpose(1, 2)
    vf = qkv[:, :, 2 * C:].float().view(B, S, H, D).transpose(1, 2)
    p = torch.softmax(qf @ kf.transpose(-1, -2) / D ** 0.5, dim=-1)
    ref = (p @ vf).transpose(1, 2).reshape(B, S, C)
    err = (out.float() - ref).abs().max().item()
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    print(f"257-token attention, keys >= {first_big} x9: max err {err:.4f}, rel {rel:.2e}")
    assert torch.isfinite(out.float()).all() and err < 0.05 and rel < 1e-2


@pytest.mark.parametrize("M,N,K", [(1024, 768, 3072), (300, 768, 768), (20000, 1024, 1024), (33024, 1024, 4096), (129, 512, 256)])
def test_gemm_fused_layernorm_tail(device, M, N, K):
    """Residual GEMM that also emits LayerNorm(out) as bf16 once a 128-row block is complete (gemm_kernel.cuh::ln_tail):
    `out` must be unchanged by the fusion and the bf16 rows must equal the stand-alone LayerNorm kernel bit for bit."""
    from embodied_captioning_b200 import _native as nat
    a = _rand_bf16((M, K), device, seed=11)
    w = _rand_bf16((N, K), device, K ** -0.5, seed=12)
    bias = torch.randn(N, device=device)
    res = torch.randn(M, N, device=device) * 3
    g = torch.randn(N, device=device) * 0.2 + 1.0
    b = torch.randn(N, device=device) * 0.1
    plain = ops.gemm(a, w, bias, ops.EPI_F32_RESID, residual=res)
    ref_ln = ops.layernorm(plain, g, b, out_bf16=True)
    out = torch.empty_like(plain)
    ln = torch.zeros((M, N), device=device, dtype=torch.bfloat16)
    cnt = torch.zeros(((M + 127) // 128,), device=device, dtype=torch.int32)
    for _ in range(2):                                     # twice: the counters must re-arm themselves
        ln.zero_()
        nat.check(nat.lib().ec_gemm_bf16_ln(a.data_ptr(), K, w.data_ptr(), K, M, N, K, bias.data_ptr(), res.data_ptr(), N,
                                            out.data_ptr(), N, g.data_ptr(), b.data_ptr(), ln.data_ptr(), N, cnt.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), "ec_gemm_bf16_ln")
        ops.check_watchdog(device)
        assert torch.equal(out, plain)
        assert torch.equal(ln, ref_ln)
        assert int(cnt.abs().sum()) == 0


# ------------------------------------------------------------------------------- absorbed cross-attention (decoder)
@pytest.mark.parametrize("M,heads,T,hd", [(1024, 12, 768, 64), (300, 8, 512, 64), (7, 4, 256, 64)])
def test_grouped_gemm_per_head_projections(device, M, heads, T, hd):
    """The two grouped GEMMs on either side of the absorbed cross-attention, vs plain fp32 per-head matmuls."""
    # q' = per-head W_k^T q: A [M, heads*hd], groups of T output columns, K = hd
    q = _rand_bf16((M, heads * hd), device, seed=11)
    wk_abs = _rand_bf16((heads * T, hd), device, hd ** -0.5, seed=12)
    out = ops.gemm_grouped(q, wk_abs, group_cols=T)
    ref = torch.einsum("mhd,htd->mht", q.float().view(M, heads, hd), wk_abs.float().view(heads, T, hd)).reshape(M, heads * T)
    assert (out.float() - ref).abs().max().item() < 0.05
    simt = ops.gemm_grouped(q, wk_abs, group_cols=T, epilogue=ops.EPI_F32, simt_reference=True)
    assert (simt - ref).abs().max().item() < 1e-3
    # att = per-head W_v c + b_v: A [M, heads*T], groups of hd output columns, K = T
    c = _rand_bf16((M, heads * T), device, seed=13)
    wv = _rand_bf16((heads * hd, T), device, T ** -0.5, seed=14)
    bias = torch.randn(heads * hd, device=device)
    out2 = ops.gemm_grouped(c, wv, group_cols=hd, bias=bias, epilogue=ops.EPI_F32)
    ref2 = torch.einsum("mht,hdt->mhd", c.float().view(M, heads, T), wv.float().view(heads, hd, T)).reshape(M, heads * hd) + bias
    assert (out2 - ref2).abs().max().item() < 2e-3 * max(1.0, ref2.abs().max().item())
    ops.check_watchdog(device)


@pytest.mark.parametrize("B,heads,T,n_kv", [(5, 12, 768, 255), (300, 12, 768, 255), (33, 8, 512, 255), (9, 4, 256, 255),
                                            (3, 12, 768, 31), (4, 12, 768, 64), (2, 4, 256, 1)])
def test_cross_attention_absorbed_matches_fp32(device, B, heads, T, n_kv):
    """ctx = softmax(qp xn^T) xn per crop and head (keys = values = the normalised image tokens), persistent over crops
    (B > 148 wraps the CTA loop), ragged last chunk (n_kv % 32 != 0)."""
    qp = _rand_bf16((B, heads, T), device, 3.0 * T ** -0.5, seed=21)      # logits of O(3): a peaked softmax
    xn = _rand_bf16((B, n_kv, T), device, seed=22)
    ctx = ops.cross_attention_absorbed(qp, ops.pack_image_tokens(xn.float()), n_kv)
    ops.check_watchdog(device)
    s = torch.einsum("bht,bjt->bhj", qp.float(), xn.float())
    ref = torch.einsum("bhj,bjt->bht", torch.softmax(s, dim=-1), xn.float())
    err = (ctx.float() - ref).abs().max().item()
    assert err < 0.03, err
    cos = torch.nn.functional.cosine_similarity(ctx.float().flatten(1), ref.flatten(1)).min().item()
    assert cos > 0.9999, cos


def test_cross_attention_absorbed_skips_finished_rows(device):
    B, heads, T, n_kv = 200, 12, 768, 255
    qp = _rand_bf16((B, heads, T), device, 2.0 * T ** -0.5, seed=31)
    xn = _rand_bf16((B, n_kv, T), device, seed=32)
    fin = (torch.arange(B, device=device) % 3 == 1).to(torch.uint8)
    packed = ops.pack_image_tokens(xn.float())
    full = ops.cross_attention_absorbed(qp, packed, n_kv)
    part = ops.cross_attention_absorbed(qp, packed, n_kv, finished=fin)
    ops.check_watchdog(device)
    live = fin == 0
    assert torch.equal(part[live], full[live])
    rev = ops.cross_attention_absorbed(qp, packed, n_kv, finished=fin, reverse=True)     # other traversal, same result
    assert torch.equal(rev, part)
    assert part[~live].abs().max().item() == 0.0          # untouched (the wrapper zero-fills its output)


@pytest.mark.parametrize("B,n_kv,T", [(3, 255, 768), (5, 31, 512), (2, 64, 256)])
def test_pack_image_tokens_layout_and_normalisation(device, B, n_kv, T):
    """Packed layout: 32-row chunks, rows of T + 8 elements, zero padding; normalize = LayerNorm without affine."""
    x = torch.randn((B, n_kv, T), device=device) * 3 + 1.5
    rows = ops.pack_image_tokens(x[:1]).numel() // (T + 8)          # n_kv rounded up to the kernel's chunk size
    assert rows >= n_kv and rows % 16 == 0
    for normalize in (False, True):
        p = ops.pack_image_tokens(x, normalize=normalize).view(B, rows, T + 8).float()
        ref = torch.nn.functional.layer_norm(x, (T,), eps=1e-5) if normalize else x
        assert (p[:, :n_kv, :T] - ref).abs().max().item() < 0.02 * ref.abs().max().item()
        assert p[:, :n_kv, T:].abs().max().item() == 0.0
        if rows > n_kv:
            assert p[:, n_kv:].abs().max().item() == 0.0


@pytest.mark.parametrize("M,N,K,S", [(1024, 768, 3072, 3), (1024, 768, 3072, 2), (300, 768, 3072, 4), (1000, 512, 2048, 4),
                                     (77, 256, 4096, 2)])
def test_gemm_split_k_is_deterministic_and_matches_fp32(device, M, N, K, S):
    """Residual GEMM with the K range cut into slices on adjacent CTAs (decoder FC2): equals the fp32 reference, is
    bit-identical from run to run (fixed summation order), and the tile counters re-arm themselves."""
    a = _rand_bf16((M, K), device, seed=41)
    w = _rand_bf16((N, K), device, K ** -0.5, seed=42)
    bias = torch.randn(N, device=device)
    res = torch.randn(M, N, device=device)
    ref = a.float() @ w.float().t() + bias + res
    cnt = torch.zeros((64,), device=device, dtype=torch.int32)
    outs = [ops.gemm_splitk(a, w, bias, res, split_k=S, counters=cnt) for _ in range(4)]
    ops.check_watchdog(device)
    assert int(cnt.abs().max().item()) == 0                       # re-armed
    assert (outs[0] - ref).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    plain = ops.gemm(a, w, bias, ops.EPI_F32_RESID, residual=res)
    assert (plain - outs[0]).abs().max().item() < 1e-3 * max(1.0, ref.abs().max().item())
